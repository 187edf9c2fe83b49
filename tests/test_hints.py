"""CU-split hints (fhc_submit_frames_hinted; BASELINE configs[4] "CNN-guided split").

The reference has no hook for the predictor; the hook is this repository's definition: a two-condition
patch of TEncCu::xCompressCU.  What pins it: oracle/_ref/TAppEncoderDumpHinted = the real HM 16.14
carrying exactly that patch (oracle/build_ref.py), its per-CTU dump committed as tests/golden/hinted_200x136_qp32.npz.
  CPU: the oracle's hinted mode against that fixture and a live run; the host emulation of the kernel source against
       the oracle; all-zero hints = the un-hinted search; the integrated encoder (shim + emulated ABI) byte-identical to
       the patched reference.
  GPU: the CUDA path through the C ABI against the oracle, with random hints and with the CNN predictor's decisions.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200.synth import synth_frames, write_yuv

FIXTURE = os.path.join(T.GOLDEN, "hinted_200x136_qp32.npz")


def have_hinted_ref():
    return os.path.exists(os.path.join(T.REF_DIR, "TAppEncoderDumpHinted"))


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
    return (w, h, n, qp, bd), (Y, U, V), z["hints"], refs


def test_oracle_hinted_matches_patched_reference_golden():
    (w, h, n, qp, bd), (Y, U, V), hints, refs = _fixture()
    p = T.params_from_qp(w, h, bd, qp)
    for k in range(n):
        assert T.compare_ctus(T.oracle_compress(p, Y[k], U[k], V[k], hints[k]), refs[k], "pic%d" % k) == []
    # the hints matter: the un-hinted search decides differently
    assert T.compare_ctus(T.oracle_compress(p, Y[0], U[0], V[0]), refs[0], "plain") != []


def test_hints_are_obeyed():
    """A 'do not split' CU is coded at depth 1, a 'split' CU deeper; CUs cut by the picture edge ignore their entry."""
    (w, h, n, qp, bd), (Y, U, V), hints, refs = _fixture()
    wc = (w + 63) // 64
    for k in range(n):
        for addr in range(T.n_ctus(w, h)):
            cx, cy = (addr % wc) * 64, (addr // wc) * 64
            for q in range(4):
                x, y = cx + 32 * (q & 1), cy + 32 * (q >> 1)
                if x >= w or y >= h:
                    continue
                d = refs[k][addr]["depth"][64 * q:64 * q + 64]
                inside = x + 32 <= w and y + 32 <= h
                if not inside:
                    assert d.min() >= 2 or refs[k][addr]["predMode"][64 * q] != 1
                elif refs[k][addr]["depth"][0] == 0:
                    continue                               # the 64x64 CU won: depth 1 was a losing hypothesis
                elif hints[k][addr][q] == 1:
                    assert np.all(d == 1), (k, addr, q)
                elif hints[k][addr][q] == 2:
                    assert d.min() >= 2, (k, addr, q)


def test_zero_hints_equal_plain_search():
    w, h, qp = 136, 72, 30
    Y, U, V = synth_frames(w, h, 1, sigma=10.0, seed=3)
    p = T.params_from_qp(w, h, 8, qp)
    z = np.zeros((T.n_ctus(w, h), 4), np.uint8)
    plain = T.oracle_compress(p, Y[0], U[0], V[0])
    assert T.compare_ctus(T.oracle_compress(p, Y[0], U[0], V[0], z), plain) == []
    assert T.compare_ctus(T.emu_compress(p, Y[0], U[0], V[0], z), plain) == []


@pytest.mark.parametrize("w,h,bd,qp,sigma", [(200, 136, 8, 32, 12.0), (136, 72, 10, 27, 12.0), (192, 128, 8, 22, 20.0), (264, 200, 8, 37, 4.0)])
def test_kernel_source_emulation_hinted_matches_oracle(w, h, bd, qp, sigma):
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=sigma, seed=w)
    hints = np.random.default_rng(h).integers(0, 3, (T.n_ctus(w, h), 4)).astype(np.uint8)
    p = T.params_from_qp(w, h, bd, qp)
    assert T.compare_ctus(T.emu_compress(p, Y[0], U[0], V[0], hints), T.oracle_compress(p, Y[0], U[0], V[0], hints)) == []


def test_hints_from_split_layout():
    sys.path.insert(0, os.path.join(T.ROOT, "oracle"))
    import cnn_split as cs
    w, h = 200, 136                       # 4 x 3 CTUs; last column 8 wide, last row 8 high
    pos = cs.cu_positions(w, h)
    split = (np.arange(len(pos)) % 2).astype(np.uint8)
    hints = T.hints_from_split(w, h, split)
    assert hints.shape == (12, 4)
    for i, (x, y) in enumerate(pos):
        assert hints[(y // 64) * 4 + x // 64, (x % 64) // 32 + 2 * ((y % 64) // 32)] == (2 if split[i] else 1)
    assert np.all(hints[3] == 0) and np.all(hints[8:] == 0) and (hints != 0).sum() == len(pos)
    from fasthevc_b200 import cnn
    assert np.array_equal(cnn.hints_from_split(w, h, split)[0], hints)        # the package's vectorised scatter
    for (ww, hh) in ((1920, 1080), (416, 240), (64, 64)):
        sp = np.random.default_rng(ww).integers(0, 2, (3, len(cs.cu_positions(ww, hh)))).astype(np.uint8)
        got = cnn.hints_from_split(ww, hh, sp)
        for k in range(3):
            assert np.array_equal(got[k], T.hints_from_split(ww, hh, sp[k]))


@pytest.mark.skipif(not have_hinted_ref(), reason="oracle/_ref/TAppEncoderDumpHinted not built")
def test_oracle_hinted_matches_live_patched_reference_and_integrated_encoder():
    """Live: random hints through (a) the patched reference, (b) the oracle, (c) the reference encoder with the shim on the
    emulated ABI (FHC_HINTS -> fhc_submit_frames_hinted).  (a) = (b) on every CTU field; (a) and (c) write the same
    bitstream; the decoder's MD5 SEI check passes."""
    w, h, n, qp = 136, 72, 3, 29
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=14.0, seed=9)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    hints = np.random.default_rng(1).integers(0, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
    hints[1] = 1
    hints.tofile(os.path.join(d, "hints.bin"))
    env = {"FHC_HINTS": os.path.join(d, "hints.bin")}
    ref, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, exe="TAppEncoderDumpHinted", env_extra=env, tag="_ref")
    pics = T.parse_dump(dump)
    for k, (s, ctus) in enumerate(pics):
        assert T.compare_ctus(T.oracle_compress(T.params_from_slice(s), Y[k], U[k], V[k], hints[k]), ctus, "pic%d" % k) == []
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_ref.bin"), d)
    assert ok == n and not bad
    if T.have_integrated():
        got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=T.build_emu_abi(),
                                    env_extra=env, tag="_fhc")
        assert got == ref


# ------------------------------------------------------------------ GPU ------------------------------------------------------
def _cuda_hinted(w, h, bd, sps, Y, U, V, hints):
    import fasthevc_b200 as fh
    n = Y.shape[0]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0)
    enc.submit(enc.pack_frames(Y, U, V), sps, hints=hints)
    out = [enc.fetch_picture(k) for k in range(n)]
    enc.close()
    return out


@pytest.mark.gpu
def test_cuda_hinted_matches_patched_reference_golden():
    (w, h, n, qp, bd), (Y, U, V), hints, refs = _fixture()
    sps = [T.slice_params_from_hm(T.params_from_qp(w, h, bd, qp))] * n
    got = _cuda_hinted(w, h, bd, sps, Y, U, V, hints)
    for k in range(n):
        assert T.compare_ctus(got[k], refs[k], "pic%d" % k) == []


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,bd,sigma", [(416, 240, 12, 8, 8.0), (200, 136, 7, 10, 12.0), (264, 200, 5, 8, 20.0)])
def test_cuda_hinted_matches_oracle(w, h, n, bd, sigma):
    """Random hints, mixed QPs, several pictures per block (hinted chains share a block without twin alignment)."""
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=77)
    qps = [22 + 5 * (k % 4) for k in range(n)]
    hints = np.random.default_rng(n).integers(0, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
    hints[0] = 1
    hints[n - 1] = 2
    ps = [T.params_from_qp(w, h, bd, q) for q in qps]
    got = _cuda_hinted(w, h, bd, [T.slice_params_from_hm(p) for p in ps], Y, U, V, hints)
    ref = T.oracle_compress_parallel([((w, h, bd, qps[k]), Y[k], U[k], V[k], hints[k]) for k in range(n)])
    for k in range(n):
        assert T.compare_ctus(got[k], ref[k], "pic%d" % k) == []


@pytest.mark.gpu
def test_cuda_hinted_and_plain_submits_in_flight():
    """A hinted submit queued behind a plain one (two submit slots): each equals its own oracle."""
    import fasthevc_b200 as fh
    w, h, n, qp = 200, 136, 4, 30
    Y, U, V = synth_frames(w, h, 2 * n, sigma=10.0, seed=5)
    hints = np.random.default_rng(8).integers(0, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
    p = T.params_from_qp(w, h, 8, qp)
    sps = [T.slice_params_from_hm(p)] * n
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    enc.submit(enc.pack_frames(Y[:n], U[:n], V[:n]), sps, first_poc=0)
    enc.submit(enc.pack_frames(Y[n:], U[n:], V[n:]), sps, first_poc=n, hints=hints)
    out = [enc.fetch_picture(k) for k in range(2 * n)]
    with pytest.raises(fh.FhcError):
        enc.submit(enc.pack_frames(Y[:n], U[:n], V[:n]), sps, first_poc=2 * n, hints=np.full_like(hints, 3))
    enc.close()
    for k in range(2 * n):
        ref = T.oracle_compress(p, Y[k], U[k], V[k], hints[k - n] if k >= n else None)
        assert T.compare_ctus(out[k], ref, "poc%d" % k) == []


@pytest.mark.gpu
def test_cnn_guided_encode_end_to_end():
    """BASELINE configs[4] in small: submit -> the CNN predictor on the pictures resident in the engine's input buffer
    (fhc_device_picture, no second host->device copy) -> hints -> hinted encode; equals the oracle driven by the fp32
    stand-in's decisions wherever the two predictors agree (they must on every CU with a clear margin)."""
    import fasthevc_b200 as fh
    from fasthevc_b200 import cnn
    sys.path.insert(0, os.path.join(T.ROOT, "oracle"))
    import cnn_split as cs
    w, h, n, qp = 416, 240, 6, 32
    Y, U, V = synth_frames(w, h, n, sigma=12.0)
    wts = cs.make_weights()
    p = T.params_from_qp(w, h, 8, qp)
    sps = [T.slice_params_from_hm(p)] * n
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    sp = cnn.SplitPredictor(wts)
    yuv = enc.pack_frames(Y, U, V)
    enc.submit(yuv, sps, first_poc=0)                          # plain search: also puts the pictures on the device
    dptr, nbytes = enc.device_picture_ptr(0)
    assert nbytes == yuv.nbytes
    logits, split = sp.predict_device(dptr, 1, 8, w, h, n)
    l2, s2 = sp.predict(yuv, 1, 8, w, h, n)
    assert np.array_equal(logits, l2) and np.array_equal(split, s2)
    hints = np.stack([T.hints_from_split(w, h, split[k]) for k in range(n)])
    enc.submit(yuv, sps, first_poc=n, hints=hints)
    plain = [enc.fetch_picture(k) for k in range(n)]
    guided = [enc.fetch_picture(n + k) for k in range(n)]
    enc.close()
    sp.close()
    ref_split = np.stack([cs.split_decision(cs.forward(wts, cs.gather_patches(Y[k]))) for k in range(n)])
    assert (ref_split == split).mean() >= 0.99
    ref = T.oracle_compress_parallel([((w, h, 8, qp), Y[k], U[k], V[k], hints[k]) for k in range(n)])
    for k in range(n):
        assert T.compare_ctus(guided[k], ref[k], "guided%d" % k) == []
    # the guided search codes a valid partitioning with an RD cost close to the full search's
    full, fast = sum(float(x["totalCost"].sum()) for x in plain), sum(float(x["totalCost"].sum()) for x in guided)
    assert fast >= full * 0.999
    print("RD cost guided / full: %.4f" % (fast / full))


@pytest.mark.gpu
@pytest.mark.skipif(not (have_hinted_ref() and T.have_integrated()), reason="oracle/_ref binaries did not travel")
def test_integrated_encoder_hinted_bitstream_cuda():
    """The reference encoder with the shim on the CUDA library, hints from FHC_HINTS: the same Annex-B bitstream as the
    reference carrying the hook patch; decoder MD5 SEI (OK)."""
    w, h, n, qp = 416, 240, 3, 32
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=8.0, seed=2)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    hints = np.random.default_rng(4).integers(1, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
    hints.tofile(os.path.join(d, "hints.bin"))
    env = {"FHC_HINTS": os.path.join(d, "hints.bin"), "FHC_SHIM_INPUT": os.path.join(d, "s.yuv")}
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderDumpHinted", env_extra=env, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", env_extra=env, tag="_fhc")
    assert got == ref
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad
