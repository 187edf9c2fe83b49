"""CPU suite (-m "not gpu"): the oracle against the committed golden vectors and the live reference, the kernel
source in host emulation against the oracle, leaf kernels, the C-ABI library surface, frame-group sharding."""
import ctypes as C
import glob
import hashlib
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200.synth import synth_frames, write_yuv

GOLDEN = sorted(p for p in glob.glob(os.path.join(T.GOLDEN, "*.npz")) if not os.path.basename(p).startswith(("stages_", "cnn_", "hinted_", "wpp_", "fullsize_", "deblock_")))


def load_golden(path):
    z = np.load(path)
    w, h, n, qp, bd = [int(v) for v in z["meta"]]
    sigma = float(z["sigma"][0])
    pics = []
    for k in range(n):
        rec = np.zeros(T.n_ctus(w, h), T.CTU_OUT_DT)
        for f in T.FIELDS:
            rec[f] = z["p%d_%s" % (k, f)]
        pics.append((z["slice%d" % k], rec))
    return (w, h, n, qp, bd, sigma), pics, z


def golden_params(w, h, bd, qp, sl):
    p = T.params_from_qp(w, h, bd, qp)
    # the reference's own doubles, as dumped (they must equal the restated lambda arithmetic bit for bit)
    assert p.lambda_ == sl[0] and p.sqrtLambda == sl[1] and p.distWeight[1] == sl[2] and p.distWeight[2] == sl[3]
    assert list(p.rdoqLambda) == list(sl[4:7])
    return p


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_oracle_matches_reference_golden(path):
    """Oracle == real reference (per-CTU dump of TAppEncoderDump) on every field of every CTU, bit exact."""
    (w, h, n, qp, bd, sigma), pics, _ = load_golden(path)
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma)
    for k, (sl, ref) in enumerate(pics):
        got = T.oracle_compress(golden_params(w, h, bd, qp, sl), Y[k], U[k], V[k])
        assert T.compare_ctus(got, ref, "pic%d" % k) == []


@pytest.mark.parametrize("path", GOLDEN[:3], ids=[os.path.basename(g)[:-4] for g in GOLDEN[:3]])
def test_kernel_source_emulation_matches_golden(path):
    """The CUDA kernel source compiled for the host (tests/emu_harness.cpp) == reference golden vectors."""
    (w, h, n, qp, bd, sigma), pics, _ = load_golden(path)
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma)
    for k, (sl, ref) in enumerate(pics[:1]):
        got = T.emu_compress(golden_params(w, h, bd, qp, sl), Y[k], U[k], V[k])
        assert T.compare_ctus(got, ref, "pic%d" % k) == []


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_and_decoder_md5():
    """Fresh content, not in the fixtures: oracle == reference dump; reference bitstream decodes with MD5 (OK)."""
    w, h, n, qp = 152, 88, 2, 29
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=15.0, seed=77)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    bs, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d)
    pics = T.parse_dump(dump)
    for k, (s, ctus) in enumerate(pics):
        got = T.oracle_compress(T.params_from_slice(s), Y[k], U[k], V[k])
        assert T.compare_ctus(got, ctus, "pic%d" % k) == []
    r = subprocess.run([os.path.join(T.REF_DIR, "TAppDecoder"), "-b", os.path.join(d, "out.bin"), "-o", os.path.join(d, "dec.yuv")],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("(OK)") == n
    assert open(os.path.join(d, "dec.yuv"), "rb").read() == open(os.path.join(d, "rec.yuv"), "rb").read()


def test_golden_bitstream_md5_recorded():
    for path in GOLDEN:
        z = np.load(path)
        assert z["bitstream_md5"].shape == (16,) and int(z["bitstream_len"][0]) > 0


def test_lambda_restatement():
    """hm_params_from_qp == TEncSlice::calculateLambda/setUpLambda for the cfg (values from a reference run, QP 32)."""
    p = T.params_from_qp(416, 240, 8, 32)
    assert p.lambda_ == 57.908390375799925 and p.distWeight[1] == 1.2599210498948732 and p.rdoqLambda[1] == 45.96191990016498


# ---------------------------------------------------------------- leaf kernels of the oracle
def _lib():
    lib = T.oracle()
    lib.hm_fwd_transform.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.hm_inv_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.hm_satd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.hm_satd.restype = C.c_uint32
    lib.hm_predict.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return lib


@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_transform_matches_float_dct(n):
    """The integer transform is the scaled DCT-II: coefficient ~ 2^(6 - shift budget) * orthonormal DCT (HEVC spec 8.6.4)."""
    lib = _lib()
    rng = np.random.default_rng(n)
    resi = rng.integers(-255, 256, (n, n)).astype(np.int16)
    coef = np.zeros((n, n), np.int32)
    lib.hm_fwd_transform(resi.ctypes.data, n, coef.ctypes.data, n, 0, 8)
    k = np.arange(n)[:, None]
    x = np.arange(n)[None, :]
    D = np.cos(np.pi * k * (2 * x + 1) / (2 * n)) * np.sqrt(2.0 / n)
    D[0] /= np.sqrt(2.0)
    ref = D @ resi.astype(np.float64) @ D.T
    scale = 2.0 ** (15 - 8 - np.log2(n))          # transform shift of HM (TComChromaFormat.h:169)
    # HEVC's matrix entries deviate from the exact cosines by up to ~4 % (e.g. 36 vs 34.6), hence the loose bound
    assert np.max(np.abs(coef - ref * scale)) <= 4.0 + 0.06 * np.max(np.abs(ref * scale))
    back = np.zeros((n, n), np.int16)
    # forward->inverse without quantisation reproduces the residual up to the transform's rounding
    lib.hm_inv_transform(coef.ctypes.data, back.ctypes.data, n, n, 0, 8)
    assert np.max(np.abs(back.astype(int) - resi)) <= 4


def test_dst_roundtrip_and_satd_known_answers():
    lib = _lib()
    resi = (np.arange(16).reshape(4, 4) * 7 - 50).astype(np.int16)
    coef = np.zeros((4, 4), np.int32)
    back = np.zeros((4, 4), np.int16)
    lib.hm_fwd_transform(resi.ctypes.data, 4, coef.ctypes.data, 4, 1, 8)
    lib.hm_inv_transform(coef.ctypes.data, back.ctypes.data, 4, 4, 1, 8)
    assert np.max(np.abs(back.astype(int) - resi)) <= 1
    a = np.zeros((8, 8), np.int16)
    b = np.zeros((8, 8), np.int16)
    assert lib.hm_satd(a.ctypes.data, 8, b.ctypes.data, 8, 8, 8) == 0
    a[:] = 4                                         # constant difference 4: only the DC Hadamard term, 64*4 -> (256+2)>>2
    assert lib.hm_satd(a.ctypes.data, 8, b.ctypes.data, 8, 8, 8) == 64
    a4 = np.full((4, 4), 3, np.int16)
    b4 = np.zeros((4, 4), np.int16)
    assert lib.hm_satd(a4.ctypes.data, 4, b4.ctypes.data, 4, 4, 8) == (48 + 1) >> 1


def test_prediction_known_answers():
    lib = _lib()
    n = 8
    ref = np.full(4 * n + 1, 100, np.int16)
    dst = np.zeros((n, n), np.int16)
    for mode in (0, 1, 10, 26, 2, 18, 34):          # flat references -> flat prediction for every mode
        lib.hm_predict(mode, n, 1, 8, ref.ctypes.data, ref.ctypes.data, 0, dst.ctypes.data, n)
        assert np.all(dst == 100), mode
    ref = np.arange(4 * n + 1).astype(np.int16)      # index 2n is the corner; above[x] = ref[2n+1+x], left[y] = ref[2n-1-y]
    lib.hm_predict(26, n, 0, 8, ref.ctypes.data, ref.ctypes.data, 0, dst.ctypes.data, n)   # chroma: pure vertical copy, no edge filter
    assert np.all(dst == ref[2 * n + 1:3 * n + 1][None, :])
    lib.hm_predict(10, n, 0, 8, ref.ctypes.data, ref.ctypes.data, 0, dst.ctypes.data, n)   # pure horizontal
    assert np.all(dst == ref[2 * n - 1 - np.arange(n)][:, None])


# ---------------------------------------------------------------- C-ABI surface (no compute without a GPU)
def test_cabi_exports_every_declared_symbol():
    import fasthevc_b200 as fh
    hdr = open(os.path.join(T.ROOT, "include", "fasthevc_cuda.h")).read()
    declared = set(re.findall(r"\b(fhc_[a-z_0-9]+)\s*\(", hdr))
    assert declared >= {"fhc_create", "fhc_submit_frames", "fhc_fetch_ctu", "fhc_destroy"}
    fh.build_library()
    lib = fh.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert C.sizeof(fh.SliceParams) == 64 and fh.CTU_RESULT_DTYPE.itemsize == 27672


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly (no oracle / emulation behind the API)."""
    import torch
    import fasthevc_b200 as fh
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fh.FhcError):
        fh.Encoder(64, 64, 8, max_frames=1)
    src = open(os.path.join(T.ROOT, "fasthevc_b200", "api.py")).read() + open(os.path.join(T.ROOT, "fasthevc_b200", "__init__.py")).read() + open(os.path.join(T.ROOT, "fasthevc_b200", "cnn.py")).read()
    assert "oracle" not in src.replace("no oracle", "") and "emu" not in src


def test_rejects_unsupported_configuration():
    import fasthevc_b200 as fh
    lib = fh.load_library()
    sq = fh.SeqParams(64, 64, 8, 1, 64, 4, 5, 2, 3, 1, 1, 1, 1, 1, 1, 1, 1)
    sq.chromaFormat = 2
    ctx = C.c_void_p()
    assert lib.fhc_create(C.byref(sq), 0, C.byref(ctx)) == -1
    sq.chromaFormat = 1
    sq.width = 65
    assert lib.fhc_create(C.byref(sq), 0, C.byref(ctx)) == -1


# ---------------------------------------------------------------- multi-GPU host logic on gloo (world size 2)
def test_frame_groups_partition():
    from fasthevc_b200.sharding import frame_group, splice_in_poc_order
    for n, ws in ((512, 8), (64, 4), (7, 2), (3, 4)):
        groups = [frame_group(n, r, ws) for r in range(ws)]
        assert groups[0][0] == 0 and groups[-1][1] == n
        assert all(groups[i][1] == groups[i + 1][0] for i in range(ws - 1))
    assert splice_in_poc_order([(2, ["c", "d"]), (0, ["a", "b"])]) == ["a", "b", "c", "d"]
    with pytest.raises(ValueError):
        splice_in_poc_order([(3, ["x"]), (0, ["a"])])


def test_two_rank_sharding_gloo():
    """Two gloo ranks each 'encode' their frame group (oracle stands in for the device), rank 0 splices in POC order and
    gets exactly the single-process result: no collective is needed on the data path, only the final gather."""
    script = os.path.join(T.ROOT, "tests", "gloo_shard_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29713")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29713", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARD_OK" in r.stdout


def read_oracle_trace(path):
    """Parses the oracle's binary stage trace (oracle/hm_intra.c, hm_intra_set_trace)."""
    import zlib
    raw = np.fromfile(path, dtype="<i4")
    tags, rmd, p2, tu, full = [], [], [], [], []
    i = 0
    while i < len(raw):
        tag, a, b, _ = (int(v) for v in raw[i:i + 4])
        i += 4
        tags.append(tag)
        if tag == ord("R"):
            rmd.append([a, raw[i - 2], raw[i - 1]])
        elif tag == ord("P"):
            p2.append([a, float(raw[i:i + 2].view("<f8")[0])])
            i += 2
        elif tag == ord("T"):
            n2 = a * a
            blocks = [raw[i + k * n2:i + (k + 1) * n2] for k in range(5)]
            i += 5 * n2
            coded = bool(np.any(blocks[2]))
            tu.append([a, b] + [zlib.crc32(np.ascontiguousarray(x).tobytes()) for x in blocks[:3]] +
                      [zlib.crc32(np.ascontiguousarray(x).tobytes()) if coded else 0 for x in blocks[3:]])
            if len(full) < 6:
                full.append([x.reshape(a, a).copy() for x in blocks[:3]])
        else:
            raise AssertionError("bad trace tag %d at word %d" % (tag, i - 4))
    return np.array(tags, np.uint8), np.array(rmd, np.int64), np.array(p2, np.float64), np.array(tu, np.int64), full


def test_oracle_stage_trace_matches_reference():
    """Per-stage pin of the oracle: for one CTU, every first-pass SATD / mode-bit pair, every candidate's full-RD cost and
    every TU's residual, transform output, RDOQ output, dequantiser and inverse-transform output - in the order the
    reference evaluates them - equal what the real reference prints with its own debug switches
    (tests/golden/stages_64x64_qp32.npz, made by tests/golden/make_stage_golden.py from oracle/_ref/TAppEncoderStages)."""
    g = np.load(os.path.join(T.GOLDEN, "stages_64x64_qp32.npz"))
    w, h, qp, bd, seed = (int(v) for v in g["params"])
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=float(g["sigma"][0]), seed=seed)
    lib = T.oracle()
    lib.hm_intra_set_trace.argtypes = [C.c_char_p]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "trace.bin")
        lib.hm_intra_set_trace(path.encode())
        try:
            T.oracle_compress(T.params_from_qp(w, h, bd, qp), Y[0], U[0], V[0])
        finally:
            lib.hm_intra_set_trace(None)
        tags, rmd, p2, tu, full = read_oracle_trace(path)
    assert len(tags) == len(g["tags"]) and np.array_equal(tags, g["tags"]), "the oracle evaluates stages in a different order than the reference"
    assert np.array_equal(rmd, g["rmd"])                                    # mode, SATD, mode bits: exact
    assert np.array_equal(p2[:, 0], g["pass2"][:, 0])
    assert np.allclose(p2[:, 1], g["pass2"][:, 1], rtol=2e-5, atol=0)       # the reference prints six significant digits
    assert np.array_equal(tu, g["tu"])                                      # size, component, five CRC-32s per TU: exact
    for k, blocks in enumerate(full):
        for key, blk in zip(("resi", "tr", "q"), blocks):
            assert np.array_equal(blk, g["full%d_%s" % (k, key)]), (k, key)


def test_twin_alignment_schedule_depends_on_geometry_only():
    """The two chains of a block wait for each other at every twin_align call of their same-role warps, so the number of
    calls per role and CTU must not depend on the picture content - a data-dependent call would deadlock the block. The
    emulation counts the calls: a smooth picture at QP 37 and a noisy one at QP 22 (different split decisions, candidate
    counts, transform-skip choices, cbf patterns) of the same size must agree role by role."""
    import ctypes as C
    emu = C.CDLL(T.build_emu())
    emu.fhc_emu_align_counts.argtypes = [C.c_void_p, C.c_int]
    w, h = 200, 136                                      # boundary CTUs in both directions
    counts = []
    for qp, sigma, seed in ((37, 1.5, 3), (22, 28.0, 4), (30, 9.0, 5)):
        Y, U, V = synth_frames(w, h, 1, sigma=sigma, seed=seed)
        out = (C.c_ulonglong * 5)()
        emu.fhc_emu_align_counts(out, 1)
        T.emu_compress(T.params_from_qp(w, h, 8, qp), Y[0], U[0], V[0])
        emu.fhc_emu_align_counts(out, 1)
        counts.append(list(out))
    assert counts[0] == counts[1] == counts[2], counts
    assert all(c > 0 for c in counts[0])


def test_scheduler_lone_picture_model():
    """Host side of the row-wise scheduler (fasthevc_cuda.cu lone_pictures): which pictures run alone in their block.
    Twin builds only: a lone chain is ~1.7x faster than a chain that shares its block, and the slowest picture ends the
    kernel - few pictures all run alone, the bench mix (64 frames x 4 QPs on 148 blocks) keeps its 64 most expensive
    pictures alone, a submit that fills every slot with equal pictures shares everything."""
    import ctypes as C
    import fasthevc_b200.api as api
    lib = C.CDLL(api.library_path())
    lib.fhc_debug_lone_pictures.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]

    def lone(qps, blocks=148):
        a = (C.c_int * len(qps))(*qps)
        return lib.fhc_debug_lone_pictures(a, len(qps), blocks)

    assert lone([32]) == 1 and lone([32, 32]) == 2 and lone([22, 37]) in (1, 2) and lone([32] * 148) == 148
    assert lone([q for q in (22, 27, 32, 37) for _ in range(64)]) == 64
    assert lone([32] * 296) == 0 and lone([32] * 1000) == 0
    assert lone([32] * 10, blocks=4) in (0, 1, 2)


def test_kernel_shared_memory_budget():
    """Two picture chains share a thread block (one block per SM): twice the chain's shared memory (Blk + role arenas,
    computed by the same block_footprint() the launch uses) plus the 1 KB the system reserves per block must fit the
    227 KB a B200 block may use, and should leave at least a third of the SM's 256 KB array to the L1 that caches the
    chains' global scratch.  Checked where the build happens."""
    import ctypes as C
    emu = C.CDLL(T.build_emu())
    smem, glob = emu.fhc_emu_shared_bytes(), emu.fhc_emu_global_bytes()
    assert 2 * smem + 1024 <= 227 * 1024, smem
    assert 2 * smem <= 170 * 1024, smem
    assert glob < 512 * 1024, glob                      # per-chain global workspace: 296 chains stay inside the 126 MB L2


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC not built")
@pytest.mark.parametrize("w,h,n,qp,bd,sigma", [(200, 136, 2, 27, 8, 12.0), (136, 72, 2, 32, 10, 10.0)])
def test_integrated_encoder_bitstream_is_byte_identical_emulated(w, h, n, qp, bd, sigma):
    """The reference encoder with TEncCu::compressCtu wired to the TLibEncoderCUDA shim (fasthevc_b200/shim), the
    shim's C ABI served by the host emulation of the kernel source: the Annex-B bitstream and the reconstruction are
    byte-identical to the stock reference's, and the decoder's MD5 SEI check passes.  Proves that what
    fhc_fetch_ctu returns is everything the unchanged host needs (the GPU suite repeats this with the CUDA library)."""
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=11)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, bit_depth=bd, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, bit_depth=bd, dump=False, exe="TAppEncoderFHC",
                                lib_dir=T.build_emu_abi(), tag="_fhc")
    assert got == ref
    assert open(os.path.join(d, "rec_fhc.yuv"), "rb").read() == open(os.path.join(d, "rec_ref.yuv"), "rb").read()
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_fhc.bin"), d)
    assert ok == n and not bad


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC not built")
def test_integrated_encoder_read_ahead_emulated():
    """The shim's read-ahead (FHC_SHIM_INPUT): 7 pictures in batches of 3 + 3 + 1 read from the input file ahead of the
    host loop - same bitstream; and with a WRONG file behind FHC_SHIM_INPUT the shim notices that the samples differ
    from what the host hands it, re-submits, and the bitstream is still the reference's."""
    w, h, n, qp = 136, 72, 7, 30
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=9.0, seed=5)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    write_yuv(os.path.join(d, "wrong.yuv"), Y[::-1], U[::-1], V[::-1])
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    for name in ("s.yuv", "wrong.yuv"):
        got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=T.build_emu_abi(),
                                    tag="_fhc", env_extra={"FHC_SHIM_INPUT": os.path.join(d, name)})
        assert got == ref, name


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC not built")
def test_integrated_encoder_keeps_two_submits_in_flight_emulated():
    """the shim queues the batch after the one the host is consuming (two submits in flight, pipelined
    option set), so the device searches batch k+1 while the host runs batch k's per-picture tail; same bitstream."""
    w, h, n, qp = 136, 72, 7, 33
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=9.0, seed=6)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    log = os.path.join(d, "calls.log")
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=T.build_emu_abi(), tag="_fhc",
                                env_extra={"FHC_SHIM_INPUT": os.path.join(d, "s.yuv"), "FHC_SHIM_LOOKAHEAD": "3", "FHC_EMU_LOG": log})
    assert got == ref
    calls = [l.split() for l in open(log).read().splitlines()]
    # [0..2] and [3..5] go out before the first fetch; [6] when the host reaches picture 3; every picture is fetched once
    assert calls[:3] == [["submit", "0", "3"], ["submit", "3", "3"], ["fetch", "0", "1"]]
    assert [c for c in calls if c[0] == "submit"] == [["submit", "0", "3"], ["submit", "3", "3"], ["submit", "6", "1"]]
    assert calls.index(["submit", "6", "1"]) == calls.index(["fetch", "3", "1"]) - 1
    assert [c[1] for c in calls if c[0] == "fetch"] == [str(k) for k in range(n)]
    # the deblocking filter of every picture ran through the boundary too (TEncGOP.cpp:1619 -> fhc_deblock_pictures)
    assert [c for c in calls if c[0] == "deblock"] == [["deblock", str(k), "1"] for k in range(n)]


@pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC not built")
def test_integrated_encoder_on_replayed_oracle_records():
    """tests/replay_abi.cpp (the measurement aid of tools/host_tail.py: a libfasthevc_cuda.so that serves precomputed CTU
    records): the integrated encoder fed with the ORACLE's records writes the stock reference's bitstream - the boundary
    carries everything the host needs, whoever computes the records."""
    w, h, n, qp = 136, 72, 3, 31
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=9.0, seed=8)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    lib_dir = os.path.join(T.ROOT, "tests", "_build", "replay")
    os.makedirs(lib_dir, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-o", os.path.join(lib_dir, "libfasthevc_cuda.so"),
                           os.path.join(T.ROOT, "tests", "replay_abi.cpp")])
    p = T.params_from_qp(w, h, 8, qp)
    np.concatenate([T.oracle_compress(p, Y[k], U[k], V[k]) for k in range(n)]).tofile(os.path.join(d, "records.bin"))
    ref, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, tag="_ref")
    got, _, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=lib_dir, tag="_fhc",
                                env_extra={"FHC_REPLAY": os.path.join(d, "records.bin")})
    assert got == ref


def test_kernel_source_emulation_4k_width_main10():
    """BASELINE configs[3] geometry in one dimension: a 3840-wide Main10 strip (60 CTUs per row, two full rows + the
    8-line bottom row) through the host emulation of the kernel source against the oracle - row-to-row neighbour and
    above-right availability at the 4K row length, 10-bit scaling of SATD / SSE / transform shifts."""
    w, h, qp, bd = 3840, 136, 27, 10
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=8.0, seed=31)
    p = T.params_from_qp(w, h, bd, qp)
    got = T.emu_compress(p, Y[0], U[0], V[0])
    ref = T.oracle_compress(p, Y[0], U[0], V[0])
    assert len(ref) == 180
    assert T.compare_ctus(got, ref, "4k strip") == []
