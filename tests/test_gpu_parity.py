"""GPU suite (-m gpu): the CUDA path, called through the C ABI (libfasthevc_cuda.so), against the oracle, the
committed golden vectors of the real reference, and - when oracle/_ref travelled to the box - the reference itself."""
import glob
import os
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200.synth import synth_frames, write_yuv

pytestmark = pytest.mark.gpu
GOLDEN = sorted(p for p in glob.glob(os.path.join(T.GOLDEN, "*.npz")) if not os.path.basename(p).startswith(("stages_", "cnn_", "hinted_", "wpp_", "fullsize_", "deblock_")))


def cuda_compress(w, h, bd, sps, Y, U, V):
    """All pictures in one submit through fhc_submit_frames / fhc_fetch_picture."""
    import fasthevc_b200 as fh
    n = Y.shape[0]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0)
    enc.submit(enc.pack_frames(Y, U, V), sps)
    out = [enc.fetch_picture(k) for k in range(n)]
    one = enc.fetch_ctu(n - 1, enc.num_ctus - 1)
    assert one.tobytes() == out[-1][-1].tobytes()
    enc.close()
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_cuda_matches_reference_golden(path):
    z = np.load(path)
    w, h, n, qp, bd = [int(v) for v in z["meta"]]
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=float(z["sigma"][0]))
    sps = [T.slice_params_from_hm(T.params_from_qp(w, h, bd, qp)) for _ in range(n)]
    got = cuda_compress(w, h, bd, sps, Y, U, V)
    for k in range(n):
        ref = np.zeros(T.n_ctus(w, h), T.CTU_OUT_DT)
        for f in T.FIELDS:
            ref[f] = z["p%d_%s" % (k, f)]
        assert T.compare_ctus(got[k], ref, "pic%d" % k) == []


@pytest.mark.parametrize("w,h,n,qp,bd,sigma", [
    (416, 240, 2, 32, 8, 6.0),        # BASELINE configs[0] geometry (boundary CTUs: 32-wide last column, 48-high last row)
    (416, 240, 1, 22, 8, 20.0),       # busy content: transform skip, NxN, deep RQT
    (264, 200, 1, 37, 8, 12.0),
    (200, 136, 2, 27, 10, 12.0),      # Main10
    (64, 64, 1, 30, 8, 8.0),          # single CTU
    (8, 8, 1, 32, 8, 8.0),            # smallest picture: one 8x8 CU, every larger CU is a forced split
])
def test_cuda_matches_oracle(w, h, n, qp, bd, sigma):
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=4321)
    p = T.params_from_qp(w, h, bd, qp)
    got = cuda_compress(w, h, bd, [T.slice_params_from_hm(p)] * n, Y, U, V)
    for k in range(n):
        ref = T.oracle_compress(p, Y[k], U[k], V[k])
        assert T.compare_ctus(got[k], ref, "pic%d" % k) == []


def test_cuda_mixed_qp_batch_is_independent():
    """Pictures of one submit are independent jobs: per-picture QP, and a picture's result does not depend on its batch."""
    w, h = 136, 72
    Y, U, V = synth_frames(w, h, 4, sigma=10.0)
    qps = [22, 27, 32, 37]
    ps = [T.params_from_qp(w, h, 8, q) for q in qps]
    got = cuda_compress(w, h, 8, [T.slice_params_from_hm(p) for p in ps], Y, U, V)
    for k in range(4):
        assert T.compare_ctus(got[k], T.oracle_compress(ps[k], Y[k], U[k], V[k]), "pic%d" % k) == []
    alone = cuda_compress(w, h, 8, [T.slice_params_from_hm(ps[2])], Y[2:3], U[2:3], V[2:3])
    assert alone[0].tobytes() == got[2].tobytes()


@pytest.mark.parametrize("n", [19, 37])
def test_cuda_many_pictures_mixed_qp(n):
    """One submit of n small pictures with different contents and QPs: every picture runs on its own thread block
    (role warps per CU depth, lane groups per candidate) and must equal the oracle."""
    w, h = 136, 72
    Y, U, V = synth_frames(w, h, n, sigma=11.0, seed=7)
    qps = [22, 37, 27, 32]
    ps = [T.params_from_qp(w, h, 8, qps[k % 4]) for k in range(n)]
    got = cuda_compress(w, h, 8, [T.slice_params_from_hm(p) for p in ps], Y, U, V)
    for k in range(n):
        assert T.compare_ctus(got[k], T.oracle_compress(ps[k], Y[k], U[k], V[k]), "pic%d" % k) == []


def test_cuda_more_pictures_than_resident_blocks():
    """More pictures than the persistent kernel keeps in flight: blocks pull further pictures from the queue."""
    import fasthevc_b200 as fh
    w, h, qp = 64, 64, 34
    n = fh.pictures_in_flight(0) + 40
    Yd, Ud, Vd = synth_frames(w, h, 6, sigma=9.0, seed=3)
    idx = np.arange(n) % 6
    p = T.params_from_qp(w, h, 8, qp)
    got = cuda_compress(w, h, 8, [T.slice_params_from_hm(p)] * n, Yd[idx], Ud[idx], Vd[idx])
    ref = [T.oracle_compress(p, Yd[k], Ud[k], Vd[k]) for k in range(6)]
    for k in range(n):
        assert got[k].tobytes() == got[int(idx[k])].tobytes()
    for k in range(6):
        assert T.compare_ctus(got[k], ref[k], "pic%d" % k) == []


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref did not travel")
def test_cuda_matches_live_reference():
    """Same input through the real reference (TAppEncoderDump, host CPU) and through the CUDA path."""
    w, h, n, qp = 328, 184, 1, 30
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=14.0, seed=99)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    _, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d)
    pics = T.parse_dump(dump)
    got = cuda_compress(w, h, 8, [T.slice_params_from_hm(T.params_from_slice(pics[0][0]))], Y, U, V)
    assert T.compare_ctus(got[0], pics[0][1], "pic0") == []


def test_full_size_1080p_all_qps_equal_oracle():
    """BASELINE configs[1] geometry and QPs: one 1920x1080 picture at each of QP 22 / 27 / 32 / 37 in one submit (four
    picture chains side by side); every field of all 4 x 510 CTUs must equal the oracle (its records for these pictures
    are committed as per-CTU CRCs, tests/golden/fullsize_1080p.npz; a mismatch re-runs the oracle to name the field)."""
    w, h, bd, sigma, seed, qps, _ = T.FULLSIZE_CASES["1080p"]
    Y, U, V = T.fullsize_frames("1080p")
    ps = [T.params_from_qp(w, h, 8, q) for q in qps]
    got = cuda_compress(w, h, 8, [T.slice_params_from_hm(p) for p in ps], Y, U, V)
    assert T.check_fullsize("1080p", got) == []
    for k in range(4):
        assert len(got[k]) == 510
        # size-independent properties: every in-picture partition coded intra, partitions below the picture untouched
        assert np.all(got[k]["predMode"][:480] == 1)
        assert all(np.any(c["predMode"] == 2) and np.any(c["predMode"] == 1) for c in got[k][480:])


def test_full_size_2160p_main10_equals_oracle():
    """BASELINE configs[3] geometry: one 3840x2160 10-bit picture (2,040 CTUs, 60 per row, 48-line bottom row), QP 27,
    encoder_intra_main10.cfg tool set; every field of every CTU equals the oracle (tests/golden/fullsize_2160p_main10.npz)."""
    w, h, bd, sigma, seed, qps, _ = T.FULLSIZE_CASES["2160p_main10"]
    Y, U, V = T.fullsize_frames("2160p_main10")
    p = T.params_from_qp(w, h, bd, qps[0])
    got = cuda_compress(w, h, bd, [T.slice_params_from_hm(p)], Y, U, V)
    assert len(got[0]) == 2040
    assert T.check_fullsize("2160p_main10", got) == []


def test_two_submits_in_flight_and_per_picture_completion():
    """The ABI accepts a second submit while the first one runs (its copy overlaps the kernel), fetch waits for its own
    picture only, poll never blocks, and both submits' results stay fetchable and exact."""
    import fasthevc_b200 as fh
    w, h = 136, 72
    Y, U, V = synth_frames(w, h, 6, sigma=9.0, seed=17)
    ps = [T.params_from_qp(w, h, 8, q) for q in (37, 22, 32, 27, 30, 35)]
    sps = [T.slice_params_from_hm(p) for p in ps]
    enc = fh.Encoder(w, h, 8, max_frames=3, device=0)
    enc.submit(enc.pack_frames(Y[:3], U[:3], V[:3]), sps[:3], first_poc=0)
    enc.submit(enc.pack_frames(Y[3:], U[3:], V[3:]), sps[3:], first_poc=3)
    assert enc.poll_picture(5) in (False, True)
    out = {}
    for poc in (4, 0, 5, 2, 1, 3):                      # any order, both submits
        out[poc] = enc.fetch_picture(poc)
        assert enc.poll_picture(poc) is True
    enc.close()
    for k in range(6):
        assert T.compare_ctus(out[k], T.oracle_compress(ps[k], Y[k], U[k], V[k]), "poc%d" % k) == []


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC did not travel")
def test_integrated_encoder_bitstream_is_byte_identical_cuda():
    """BASELINE configs[0] geometry (416x240, QP 32, boundary CTUs) through the reference encoder whose
    TEncCu::compressCtu calls the TLibEncoderCUDA shim, linked against libfasthevc_cuda.so: the whole Annex-B bitstream
    is byte-identical to the stock reference encoder's and decodes with matching MD5 SEI."""
    w, h, n, qp = 416, 240, 2, 32
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=6.0)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", tag="_fhc")
    assert got == ref
    assert open(os.path.join(d, "rec_fhc.yuv"), "rb").read() == open(os.path.join(d, "rec_ref.yuv"), "rb").read()
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC did not travel")
def test_integrated_encoder_1080p_bitstream_is_byte_identical_cuda():
    """The metric's geometry end to end: 1920x1080, 2 frames, QP 32 through the reference encoder with the
    TLibEncoderCUDA shim on the CUDA library - byte-identical Annex-B bitstream, decoder MD5 SEI (OK)."""
    w, h, n, qp = 1920, 1080, 2, 32
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=6.0)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", tag="_fhc",
                                env_extra={"FHC_SHIM_INPUT": os.path.join(d, "s.yuv")})
    assert got == ref
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC did not travel")
def test_integrated_encoder_read_ahead_cuda():
    """The shim reading ahead (FHC_SHIM_INPUT): 24 pictures of 200x136 submitted to the device in one batch, encoded
    side by side (one thread block each), consumed by the unchanged host loop picture by picture: byte-identical bitstream."""
    w, h, n, qp = 200, 136, 24, 34
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=8.0, seed=21)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", tag="_fhc",
                                env_extra={"FHC_SHIM_INPUT": os.path.join(d, "s.yuv")})
    assert got == ref
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad


def test_pipelined_submits_equal_oracle_and_plain():
    """FHC_OPT_PIPELINED_SUBMITS: two submits queued back to back, their kernels on separate streams with separate scratch,
    blocks leaving early when the rest fits into the others.  Small case against the oracle; a case with more pictures than
    chain slots (2 x 320 pictures of 416x240) against the same pictures searched one submit at a time."""
    import fasthevc_b200 as fh
    w, h, n = 200, 136, 12
    Y, U, V = synth_frames(w, h, 2 * n, sigma=11.0, seed=41)
    qps = [22 + 5 * (k % 4) for k in range(2 * n)]
    ps = [T.params_from_qp(w, h, 8, q) for q in qps]
    sps = [T.slice_params_from_hm(p) for p in ps]
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    enc.set_pipelined(True)
    enc.span_begin()
    enc.submit(enc.pack_frames(Y[:n], U[:n], V[:n]), sps[:n], first_poc=0)
    enc.submit(enc.pack_frames(Y[n:], U[n:], V[n:]), sps[n:], first_poc=n)
    got = [enc.fetch_picture(k) for k in range(2 * n)]
    assert enc.span_ms() > 0
    enc.close()
    ref = T.oracle_compress_parallel([((w, h, 8, qps[k]), Y[k], U[k], V[k]) for k in range(2 * n)])
    for k in range(2 * n):
        assert T.compare_ctus(got[k], ref[k], "poc%d" % k) == []

    w, h, n = 416, 240, 320
    Y, U, V = synth_frames(w, h, 16, sigma=9.0, seed=43)
    idx = [k % 16 for k in range(n)]
    sps = [fh.slice_params_from_qp(22 + 5 * ((k // 16) % 4)) for k in range(n)]
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    yuv = enc.pack_frames(Y[idx], U[idx], V[idx])
    enc.submit(yuv, sps, first_poc=0)
    plain = [T.ctu_crcs(enc.fetch_picture(k)) for k in range(n)]        # (per-field CRCs: the record's padding bytes are never written)
    t_plain = enc.last_kernel_ms()
    enc.set_pipelined(True)
    enc.span_begin()
    enc.submit(yuv, sps, first_poc=1000)
    enc.submit(yuv, sps, first_poc=2000)
    enc.submit(yuv, sps, first_poc=3000)             # waits for the first one's slot
    for first in (2000, 3000):
        for k in range(n):
            assert np.array_equal(T.ctu_crcs(enc.fetch_picture(first + k)), plain[k]), (first, k)
    t_span = enc.span_ms()
    enc.close()
    print("320 pictures 416x240: one submit %.0f ms; three pipelined submits %.0f ms (%.0f ms each)" % (t_plain, t_span, t_span / 3))


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC did not travel")
@pytest.mark.parametrize("wpp", [False, True], ids=["plain", "wavefront"])
def test_integrated_encoder_1080p_pipeline_cuda(wpp):
    """SURVEY 8f-2 end to end: the reference encoder with the shim on the CUDA library, 24 pictures of 1920x1080 in
    batches of 12, two submits in flight (the device searches batch k+1 while the unchanged single-threaded host runs
    batch k's deblocking / SAO / final CABAC / MD5).  Every picture decodes with matching MD5 SEI; the wall time is
    printed (the stock encoder needs ~5.6 s per picture on one core at this QP)."""
    import time
    w, h, n, qp = 1920, 1080, 24, 32
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, 6, sigma=6.0)
    idx = [k % 6 for k in range(n)]
    write_yuv(os.path.join(d, "s.yuv"), Y[idx], U[idx], V[idx])
    t0 = time.perf_counter()
    got, _, out = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", tag="_fhc",
                                  env_extra={"FHC_SHIM_INPUT": os.path.join(d, "s.yuv"), "FHC_SHIM_LOOKAHEAD": "12"},
                                  extra=["--WaveFrontSynchro=1"] if wpp else [])
    wall = time.perf_counter() - t0
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad
    print("integrated encoder, %s: %d pictures 1080p QP %d in %.1f s = %.2f fps (whole bitstream, %d bytes)"
          % ("wavefront" if wpp else "plain", n, qp, wall, n / wall, len(got)))
