"""Deblocking of finished pictures on the device (fhc_deblock_pictures; SURVEY.md 8f rank 1, first stage of the
per-picture tail).  Reference: TComLoopFilter::loopFilterPic (TComLoopFilter.cpp:130-954), called from TEncGOP.cpp:1619.
Pinned against the reference encoder's reconstruction written with --SAO=0 (committed fixtures + a live run), and -
through the integrated encoder, whose deblocking call goes to the device - by whole-bitstream identity with SAO on
(tests/test_cpu_oracle.py, test_gpu_parity.py: every integrated-encoder test).
  CPU: the C oracle (oracle/hm_deblock.c) against the fixtures and a live reference run; the host build of the kernel
       source behind the emulated ABI against the oracle.
  GPU: the CUDA kernels through the C ABI against the oracle: small pictures with boundary CTUs, Main10, many pictures in
       one call, the metric's 1080p geometry, and while the mode-decision kernel of a later submit is running.
"""
import ctypes as C
import glob
import os
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200.synth import synth_frames, write_yuv

FIXTURES = sorted(glob.glob(os.path.join(T.GOLDEN, "deblock_*.npz")))


def _records(z):
    w, h, n, qp, bd = [int(v) for v in z["meta"]]
    recs = np.zeros(T.n_ctus(w, h), T.CTU_OUT_DT)
    for f in ("depth", "trIdx", "recoY", "recoCb", "recoCr"):
        recs[f] = z[f]
    return (w, h, qp, bd), recs


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_deblock_oracle_matches_reference_golden(path):
    z = np.load(path)
    (w, h, qp, bd), recs = _records(z)
    got = T.oracle_deblock(w, h, bd, qp, recs)
    want = T.planes_of(z["deblocked"], w, h)
    for c in range(3):
        assert np.array_equal(got[c], want[c]), c
    # the filter does something: the deblocked luma differs from the records' reconstruction
    pre = np.zeros((h, w), np.int16)
    wc = (w + 63) // 64
    for a in range(len(recs)):
        x0, y0 = (a % wc) * 64, (a // wc) * 64
        blk = recs[a]["recoY"].reshape(64, 64)
        pre[y0:y0 + 64, x0:x0 + 64] = blk[:min(64, h - y0), :min(64, w - x0)]
    assert (pre != want[0]).sum() > 50


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_deblock_oracle_matches_live_reference():
    w, h, n, qp = 264, 200, 2, 30
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=15.0, seed=12)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    _, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, extra=["--SAO=0"])
    rec = np.fromfile(os.path.join(d, "rec.yuv"), dtype=np.uint8)
    fs = w * h * 3 // 2
    for k, (s, ctus) in enumerate(T.parse_dump(dump)):
        recs = T.oracle_compress(T.params_from_slice(s), Y[k], U[k], V[k])
        got = T.oracle_deblock(w, h, 8, qp, recs)
        want = T.planes_of(rec[k * fs:(k + 1) * fs], w, h)
        for c in range(3):
            assert np.array_equal(got[c], want[c]), (k, c)


def test_kernel_source_deblock_emulated_matches_oracle():
    """The segment functions of fasthevc_b200/csrc/fhc_deblock.cuh compiled for the host (tests/emu_abi.cpp), through the
    same ABI calls the shim makes."""
    lib = C.CDLL(os.path.join(T.build_emu_abi(), "libfasthevc_cuda.so"))
    lib.fhc_create.argtypes = [C.POINTER(T.FhcSeqParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.fhc_submit_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_deblock_pictures.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.fhc_fetch_ctu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    for (w, h, bd, qp) in ((136, 72, 8, 27), (72, 136, 10, 34)):
        Y, U, V = synth_frames(w, h, 2, bit_depth=bd, sigma=14.0, seed=w)
        sq = T.FhcSeqParams(w, h, bd, 1, 64, 4, 5, 2, 3, 1, 1, 1, 1, 1, 1, 1, 2, 0)
        ctx = C.c_void_p()
        assert lib.fhc_create(C.byref(sq), 0, C.byref(ctx)) == 0
        p = T.params_from_qp(w, h, bd, qp)
        sps = (T.FhcSliceParams * 2)()
        for k in range(2):
            sp = T.slice_params_from_hm(p)
            C.memmove(C.byref(sps[k]), C.byref(sp), C.sizeof(sp))
        dt = np.uint8 if bd == 8 else np.uint16
        yuv = np.concatenate([Y.reshape(2, -1), U.reshape(2, -1), V.reshape(2, -1)], axis=1).astype(dt)
        assert lib.fhc_submit_frames(ctx, 7, 2, yuv.ctypes.data, yuv.itemsize, sps) == 0
        out = np.zeros((2, w * h * 3 // 2), dt)
        assert lib.fhc_deblock_pictures(ctx, 7, 2, out.ctypes.data, yuv.itemsize) == 0
        for k in range(2):
            recs = T.oracle_compress(p, Y[k], U[k], V[k])
            want = T.oracle_deblock(w, h, bd, qp, recs)
            got = T.planes_of(out[k], w, h)
            for c in range(3):
                assert np.array_equal(got[c], want[c]), (w, k, c)
        lib.fhc_destroy.argtypes = [C.c_void_p]
        lib.fhc_destroy(ctx)


# ------------------------------------------------------------------ GPU ------------------------------------------------------
def _check_cuda(enc, w, h, bd, qps, first, n, recs=None):
    out, ms = enc.deblock_pictures(first, n)
    for k in range(n):
        r = enc.fetch_picture(first + k) if recs is None else recs[k]
        want = T.oracle_deblock(w, h, bd, qps[k], r.astype(T.CTU_OUT_DT))
        got = T.planes_of(out[k], w, h)
        for c in range(3):
            assert np.array_equal(got[c], want[c]), (first + k, c, int((got[c] != want[c]).sum()))
    return ms


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,bd,sigma", [(416, 240, 6, 8, 8.0), (200, 136, 5, 10, 14.0), (64, 64, 2, 8, 20.0), (8, 8, 1, 8, 5.0), (264, 200, 37, 8, 10.0)])
def test_cuda_deblock_matches_oracle(w, h, n, bd, sigma):
    import fasthevc_b200 as fh
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=19)
    qps = [22 + 5 * (k % 4) for k in range(n)]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0)
    enc.submit(enc.pack_frames(Y, U, V), [fh.slice_params_from_qp(q) for q in qps], first_poc=3)
    got0 = enc.fetch_picture(3)
    assert T.compare_ctus(got0, T.oracle_compress(T.params_from_qp(w, h, bd, qps[0]), Y[0], U[0], V[0])) == []
    _check_cuda(enc, w, h, bd, qps, 3, n)
    if n > 2:
        _check_cuda(enc, w, h, bd, qps[1:], 4, 2)            # a sub-range of the submit
    with pytest.raises(fh.FhcError):
        enc.deblock_pictures(3, n + 1)
    enc.close()


@pytest.mark.gpu
def test_cuda_deblock_1080p_and_while_a_kernel_runs():
    """The metric's geometry, and the deblocking kernels running next to the persistent mode-decision kernel of a later
    submit (their blocks fit on an SM beside one of its blocks)."""
    import fasthevc_b200 as fh
    w, h, n = 1920, 1080, 4
    Y, U, V = synth_frames(w, h, n, sigma=6.0)
    qps = [37, 32, 37, 32]
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    yuv = enc.pack_frames(Y, U, V)
    sps = [fh.slice_params_from_qp(q) for q in qps]
    enc.submit(yuv, sps, first_poc=0)
    recs = [enc.fetch_picture(k) for k in range(n)]
    ms = _check_cuda(enc, w, h, 8, qps, 0, n, recs)
    enc.submit(yuv, sps, first_poc=100)                      # runs for seconds
    ms2 = _check_cuda(enc, w, h, 8, qps, 0, n, recs)         # the first submit's pictures again, meanwhile
    running = not enc.poll_picture(100 + n - 1)
    enc.fetch_picture(100 + n - 1)
    enc.close()
    gb = n * w * h * 1.5 * (2 + 1) / 1e9
    print("deblock 4 x 1080p: %.3f ms alone (%.0f GB/s of records in + pictures out), %.3f ms beside the running kernel (%s)"
          % (ms, gb / (ms / 1e3), ms2, "still running" if running else "kernel had finished"))


# ------------------------------------------------------------------ SAO statistics ------------------------------------------
def _org_planes(w, h, bd, sigma, seed=1234):
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=sigma, seed=seed)
    return Y[0], U[0], V[0]


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_sao_statistics_oracle_matches_reference_golden(path):
    """oracle/hm_sao.c (HM's line-buffer loops restated) against the reference's m_statData of the same picture."""
    z = np.load(path)
    (w, h, qp, bd), recs = _records(z)
    got = T.oracle_sao_stats(w, h, bd, T.planes_of(z["deblocked"], w, h), _org_planes(w, h, bd, float(z["sigma"][0])))
    assert np.array_equal(got["diff"], z["sao_diff"]) and np.array_equal(got["count"], z["sao_count"])
    assert got["count"].sum() > 4 * w * h               # every type counts most samples


def test_kernel_source_sao_statistics_emulated_matches_oracle():
    """The closed-form per-sample membership / class functions of fasthevc_b200/csrc/fhc_sao.cuh (what the kernel runs)
    against the oracle's restatement of HM's loops: pictures with boundary CTUs in both directions, one CTU, Main10."""
    lib = C.CDLL(os.path.join(T.build_emu_abi(), "libfasthevc_cuda.so"))
    lib.fhc_create.argtypes = [C.POINTER(T.FhcSeqParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.fhc_submit_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_deblock_pictures_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_destroy.argtypes = [C.c_void_p]
    for (w, h, bd, qp) in ((200, 136, 8, 34), (64, 64, 8, 30), (72, 136, 10, 36)):
        Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=5.0, seed=h)
        sq = T.FhcSeqParams(w, h, bd, 1, 64, 4, 5, 2, 3, 1, 1, 1, 1, 1, 1, 1, 1, 0)
        ctx = C.c_void_p()
        assert lib.fhc_create(C.byref(sq), 0, C.byref(ctx)) == 0
        p = T.params_from_qp(w, h, bd, qp)
        sp = T.slice_params_from_hm(p)
        dt = np.uint8 if bd == 8 else np.uint16
        yuv = np.concatenate([Y.reshape(1, -1), U.reshape(1, -1), V.reshape(1, -1)], axis=1).astype(dt)
        assert lib.fhc_submit_frames(ctx, 0, 1, yuv.ctypes.data, yuv.itemsize, C.byref(sp)) == 0
        out = np.zeros((1, w * h * 3 // 2), dt)
        stats = np.zeros((T.n_ctus(w, h), 3), T.SAO_STATS_DT)
        assert lib.fhc_deblock_pictures_stats(ctx, 0, 1, out.ctypes.data, yuv.itemsize, stats.ctypes.data) == 0
        lib.fhc_destroy(ctx)
        want = T.oracle_sao_stats(w, h, bd, T.planes_of(out[0], w, h), (Y[0], U[0], V[0]))
        got = T.sao_compact_to_hm(stats)
        assert np.array_equal(got["diff"], want["diff"]) and np.array_equal(got["count"], want["count"]), (w, h)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,bd,sigma", [(416, 240, 5, 8, 6.0), (200, 136, 3, 10, 5.0), (64, 64, 1, 8, 5.0), (1920, 1080, 2, 8, 6.0)])
def test_cuda_sao_statistics_match_oracle(w, h, n, bd, sigma):
    import fasthevc_b200 as fh
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=23)
    qps = [37 - 5 * (k % 2) for k in range(n)]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0, wpp=True)
    enc.submit(enc.pack_frames(Y, U, V), [fh.slice_params_from_qp(q) for q in qps])
    out, stats = enc.deblock_pictures_stats(0, n)
    out2, _ = enc.deblock_pictures(0, n)
    assert np.array_equal(out, out2)
    enc.close()
    for k in range(n):
        want = T.oracle_sao_stats(w, h, bd, T.planes_of(out[k], w, h), (Y[k], U[k], V[k]))
        got = T.sao_compact_to_hm(stats[k])
        assert np.array_equal(got["diff"], want["diff"]) and np.array_equal(got["count"], want["count"]), k
