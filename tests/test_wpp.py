"""Wavefront variant: --WaveFrontSynchro=1 = fhc_seq_params.waveFrontSynchro.

With entropy_coding_sync the reference resets the RD coder at the start of every CTU row - contexts and the
fractional-bit carry - and loads the contexts saved after the second CTU of the row above (TEncSlice.cpp:808-827,
:925-928).  The CTU rows of a picture are then independent but for the usual two-CTU lag, and the engine searches them
concurrently on different chain slots.  Results are pinned against the real reference run with that flag.
  CPU: oracle vs the committed fixture of the reference and a live run; the host emulation of the kernel source in the
       tightest interleavings the device may produce; the integrated encoder (shim + emulated ABI) byte-identical.
  GPU: the CUDA path against the oracle (small pictures, one-CTU-wide and one-row pictures, 1080p, 2160p Main10, many
       pictures, hints) and the integrated encoder on the CUDA library.
"""
import os
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200.synth import synth_frames, write_yuv

FIXTURE = os.path.join(T.GOLDEN, "wpp_264x200_qp32.npz")


def _fixture():
    z = np.load(FIXTURE)
    w, h, n, qp, bd = [int(v) for v in z["meta"]]
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=float(z["sigma"][0]))
    refs = []
    for k in range(n):
        ref = np.zeros(T.n_ctus(w, h), T.CTU_OUT_DT)
        for f in T.FIELDS:
            ref[f] = z["p%d_%s" % (k, f)]
        refs.append(ref)
    return (w, h, n, qp, bd), (Y, U, V), refs


def test_oracle_wpp_matches_reference_golden():
    (w, h, n, qp, bd), (Y, U, V), refs = _fixture()
    p = T.params_from_qp(w, h, bd, qp)
    assert T.compare_ctus(T.oracle_compress(p, Y[0], U[0], V[0], wpp=True), refs[0]) == []
    plain = T.oracle_compress(p, Y[0], U[0], V[0])
    wc = (w + 63) // 64
    # the first CTU row is the same with and without wavefronts; the rows below are not (other contexts, no bit carry)
    assert T.compare_ctus(plain[:wc], refs[0][:wc]) == []
    assert T.compare_ctus(plain, refs[0]) != []


@pytest.mark.parametrize("order", [0, 1, 2])
def test_kernel_source_emulation_wpp_matches_golden(order):
    """order 1 / 2: every row exactly two CTUs behind the row above, one CTU per call, each row on its own block, the
    rows of a time step visited top-down / bottom-up - the tightest schedule the device may produce (ping-pong line
    buffers, sync contexts, per-row progress)."""
    (w, h, n, qp, bd), (Y, U, V), refs = _fixture()
    p = T.params_from_qp(w, h, bd, qp)
    assert T.compare_ctus(T.emu_compress(p, Y[0], U[0], V[0], wpp=True, order=order), refs[0]) == []


@pytest.mark.parametrize("w,h,bd,qp", [(56, 200, 8, 30), (200, 56, 8, 27), (136, 136, 10, 32), (128, 192, 8, 37)])
def test_kernel_source_emulation_wpp_matches_oracle(w, h, bd, qp):
    """One CTU wide (no above-right CTU: rows restart from the initial contexts), one row high, Main10, exactly two CTUs wide."""
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=10.0, seed=w + h)
    p = T.params_from_qp(w, h, bd, qp)
    ref = T.oracle_compress(p, Y[0], U[0], V[0], wpp=True)
    for order in (0, 2):
        assert T.compare_ctus(T.emu_compress(p, Y[0], U[0], V[0], wpp=True, order=order), ref) == []
    hints = np.random.default_rng(w).integers(0, 3, (T.n_ctus(w, h), 4)).astype(np.uint8)
    assert T.compare_ctus(T.emu_compress(p, Y[0], U[0], V[0], hints, wpp=True, order=1), T.oracle_compress(p, Y[0], U[0], V[0], hints, wpp=True)) == []


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_oracle_wpp_matches_live_reference_and_integrated_encoder():
    w, h, n, qp = 200, 136, 2, 29
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=14.0, seed=9)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    ref, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, extra=["--WaveFrontSynchro=1"], tag="_ref")
    for k, (s, ctus) in enumerate(T.parse_dump(dump)):
        assert T.compare_ctus(T.oracle_compress(T.params_from_slice(s), Y[k], U[k], V[k], wpp=True), ctus, "pic%d" % k) == []
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_ref.bin"), d)
    assert ok == n and not bad
    if T.have_integrated():
        got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=T.build_emu_abi(),
                                    extra=["--WaveFrontSynchro=1"], tag="_fhc")
        assert got == ref


# ------------------------------------------------------------------ GPU ------------------------------------------------------
def _cuda_wpp(w, h, bd, sps, Y, U, V, hints=None):
    import fasthevc_b200 as fh
    n = Y.shape[0]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0, wpp=True)
    enc.submit(enc.pack_frames(Y, U, V), sps, hints=hints)
    out = [enc.fetch_picture(k) for k in range(n)]
    ms = enc.last_kernel_ms()
    enc.close()
    return out, ms


@pytest.mark.gpu
def test_cuda_wpp_matches_reference_golden():
    (w, h, n, qp, bd), (Y, U, V), refs = _fixture()
    got, _ = _cuda_wpp(w, h, bd, [T.slice_params_from_hm(T.params_from_qp(w, h, bd, qp))] * n, Y, U, V)
    assert T.compare_ctus(got[0], refs[0]) == []


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,bd,sigma", [(416, 240, 9, 8, 8.0), (56, 200, 3, 8, 10.0), (200, 56, 3, 8, 10.0), (264, 200, 40, 10, 12.0),
                                            (128, 192, 2, 8, 20.0)])
def test_cuda_wpp_matches_oracle(w, h, n, bd, sigma):
    """Several pictures, mixed QPs; 40 pictures x 4 rows = more rows than a small launch has chain slots for at once."""
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=31)
    qps = [22 + 5 * (k % 4) for k in range(n)]
    ps = [T.params_from_qp(w, h, bd, q) for q in qps]
    got, _ = _cuda_wpp(w, h, bd, [T.slice_params_from_hm(p) for p in ps], Y, U, V)
    ref = T.oracle_compress_parallel([((w, h, bd, qps[k]), Y[k], U[k], V[k], None, True) for k in range(n)])
    for k in range(n):
        assert T.compare_ctus(got[k], ref[k], "pic%d" % k) == []


@pytest.mark.gpu
def test_cuda_wpp_hinted_matches_oracle():
    w, h, n, qp = 416, 240, 5, 30
    Y, U, V = synth_frames(w, h, n, sigma=12.0, seed=3)
    hints = np.random.default_rng(2).integers(0, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
    p = T.params_from_qp(w, h, 8, qp)
    got, _ = _cuda_wpp(w, h, 8, [T.slice_params_from_hm(p)] * n, Y, U, V, hints)
    ref = T.oracle_compress_parallel([((w, h, 8, qp), Y[k], U[k], V[k], hints[k], True) for k in range(n)])
    for k in range(n):
        assert T.compare_ctus(got[k], ref[k], "pic%d" % k) == []


@pytest.mark.gpu
def test_cuda_wpp_full_size_1080p_and_latency():
    """The metric's geometry with wavefronts: one picture per QP, 17 rows each, all 4 x 510 CTUs against the oracle
    (tests/golden/fullsize_1080p_wpp.npz); a single 1080p picture finishes several times sooner than as one chain (that is
    what the variant is for)."""
    import fasthevc_b200 as fh
    w, h, bd, sigma, seed, qps, _ = T.FULLSIZE_CASES["1080p_wpp"]
    Y, U, V = T.fullsize_frames("1080p_wpp")
    got, ms = _cuda_wpp(w, h, 8, [fh.slice_params_from_qp(q) for q in qps], Y, U, V)
    assert T.check_fullsize("1080p_wpp", got) == []
    one, ms1 = _cuda_wpp(w, h, 8, [fh.slice_params_from_qp(qps[2])], Y[2:3], U[2:3], V[2:3])
    assert np.array_equal(T.ctu_crcs(one[0]), T.ctu_crcs(got[2]))
    print("1080p wavefront: 4 pictures %.0f ms, one QP-32 picture %.0f ms" % (ms, ms1))
    assert ms1 < 4000           # as one chain: ~7.4 s


@pytest.mark.gpu
def test_cuda_wpp_full_size_2160p_main10():
    w, h, bd, sigma, seed, qps, _ = T.FULLSIZE_CASES["2160p_main10_wpp"]
    Y, U, V = T.fullsize_frames("2160p_main10_wpp")
    got, ms = _cuda_wpp(w, h, bd, [T.slice_params_from_hm(T.params_from_qp(w, h, bd, qps[0]))], Y, U, V)
    assert T.check_fullsize("2160p_main10_wpp", got) == []
    print("2160p Main10 wavefront: one picture %.0f ms" % ms)


@pytest.mark.gpu
@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC did not travel")
def test_integrated_encoder_wpp_bitstream_cuda():
    """The reference encoder with --WaveFrontSynchro=1, compressCtu served by the CUDA library: byte-identical bitstream
    (entry points and all), decoder MD5 SEI (OK)."""
    w, h, n, qp = 416, 240, 3, 32
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=8.0, seed=2)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    env = {"FHC_SHIM_INPUT": os.path.join(d, "s.yuv")}
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, extra=["--WaveFrontSynchro=1"], tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", env_extra=env,
                                extra=["--WaveFrontSynchro=1"], tag="_fhc")
    assert got == ref
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad


def test_fullsize_crc_fixtures_present_and_sensitive():
    """The per-CTU CRC fixtures of the full-size GPU tests exist with the right shapes, and a single changed sample or
    cost bit changes a CTU's CRC."""
    for tag, (w, h, bd, sigma, seed, qps, wpp) in T.FULLSIZE_CASES.items():
        crc = np.load(os.path.join(T.GOLDEN, "fullsize_%s.npz" % tag))["crc"]
        assert crc.shape == (len(qps), T.n_ctus(w, h)) and crc.dtype == np.uint32
    (w, h, n, qp, bd), (Y, U, V), refs = _fixture()
    a = T.ctu_crcs(refs[0])
    r2 = refs[0].copy()
    r2["recoY"][3][17] ^= 1
    r2["totalCost"][5] = np.nextafter(r2["totalCost"][5], np.inf)
    b = T.ctu_crcs(r2)
    assert [i for i in range(len(a)) if a[i] != b[i]] == [3, 5]
    assert len(set(a.tolist())) == len(a)
