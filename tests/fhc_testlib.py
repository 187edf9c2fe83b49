"""Shared helpers for the test-suite: record layouts, the oracle (ctypes), the
reference binaries in oracle/_ref, synthetic content.  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# dump records written by oracle/ref_dump_hook.cpp (packed)
SLICE_DT = np.dtype([("magic", "<u4"), ("poc", "<i4"), ("qp", "<i4"), ("width", "<i4"), ("height", "<i4"),
                     ("bitDepthY", "<i4"), ("bitDepthC", "<i4"), ("lambda", "<f8"), ("sqrtLambda", "<f8"),
                     ("distWeightCb", "<f8"), ("distWeightCr", "<f8"), ("rdoqLambda", "<f8", 3)])
CTU_DUMP_DT = np.dtype([("magic", "<u4"), ("poc", "<i4"), ("ctuRsAddr", "<i4"),
                        ("depth", "u1", 256), ("partSize", "u1", 256), ("predMode", "u1", 256), ("lumaDir", "u1", 256),
                        ("chromaDir", "u1", 256), ("trIdx", "u1", 256), ("cbf", "u1", (3, 256)), ("tskip", "u1", (3, 256)),
                        ("coefY", "<i4", 4096), ("coefCb", "<i4", 1024), ("coefCr", "<i4", 1024),
                        ("recoY", "<i2", 4096), ("recoCb", "<i2", 1024), ("recoCr", "<i2", 1024),
                        ("totalCost", "<f8"), ("totalDist", "<u4"), ("totalBits", "<u4"), ("totalBins", "<u4")])
# one finished CTU as produced by the oracle and by the C-ABI (include/fasthevc_cuda.h: fhc_ctu_result)
CTU_OUT_DT = np.dtype([("depth", "u1", 256), ("partSize", "u1", 256), ("predMode", "u1", 256), ("lumaDir", "u1", 256),
                       ("chromaDir", "u1", 256), ("trIdx", "u1", 256), ("cbf", "u1", (3, 256)), ("tskip", "u1", (3, 256)),
                       ("coefY", "<i2", 4096), ("coefCb", "<i2", 1024), ("coefCr", "<i2", 1024),
                       ("recoY", "<i2", 4096), ("recoCb", "<i2", 1024), ("recoCr", "<i2", 1024),
                       ("totalCost", "<f8"), ("totalDist", "<u4"), ("totalBits", "<u4"), ("totalBins", "<u4")], align=True)
FIELDS = ["depth", "partSize", "predMode", "lumaDir", "chromaDir", "trIdx", "cbf", "tskip", "coefY", "coefCb", "coefCr",
          "recoY", "recoCb", "recoCr", "totalCost", "totalDist", "totalBits", "totalBins"]


class HmParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bitDepth", C.c_int32), ("qp", C.c_int32),
                ("lambda_", C.c_double), ("sqrtLambda", C.c_double), ("distWeight", C.c_double * 3),
                ("rdoqLambda", C.c_double * 3)]


def build_oracle():
    so3 = os.path.join(ORACLE_DIR, "libhm_sao_oracle.so")
    src3 = os.path.join(ORACLE_DIR, "hm_sao.c")
    if not os.path.exists(so3) or os.path.getmtime(so3) < os.path.getmtime(src3):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-o", so3, src3])
    so2 = os.path.join(ORACLE_DIR, "libhm_deblock_oracle.so")
    src2 = os.path.join(ORACLE_DIR, "hm_deblock.c")
    if not os.path.exists(so2) or os.path.getmtime(so2) < os.path.getmtime(src2):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so2, src2, "-lm"])
    so = os.path.join(ORACLE_DIR, "libhm_intra_oracle.so")
    src = os.path.join(ORACLE_DIR, "hm_intra.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.hm_params_from_qp.argtypes = [C.POINTER(HmParams), C.c_int, C.c_int, C.c_int, C.c_int]
        lib.hm_intra_compress_picture.argtypes = [C.POINTER(HmParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hm_intra_compress_picture.restype = C.c_int
        lib.hm_intra_compress_picture_hinted.argtypes = [C.POINTER(HmParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hm_intra_compress_picture_hinted.restype = C.c_int
        lib.hm_intra_compress_picture_ex.argtypes = [C.POINTER(HmParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.hm_intra_compress_picture_ex.restype = C.c_int
        _oracle = lib
    return _oracle


def params_from_qp(width, height, bit_depth, qp):
    p = HmParams()
    oracle().hm_params_from_qp(C.byref(p), width, height, bit_depth, qp)
    return p


def params_from_slice(rec):
    p = HmParams()
    p.width, p.height, p.bitDepth, p.qp = int(rec["width"]), int(rec["height"]), int(rec["bitDepthY"]), int(rec["qp"])
    p.lambda_, p.sqrtLambda = float(rec["lambda"]), float(rec["sqrtLambda"])
    p.distWeight[0], p.distWeight[1], p.distWeight[2] = 1.0, float(rec["distWeightCb"]), float(rec["distWeightCr"])
    for i in range(3):
        p.rdoqLambda[i] = float(rec["rdoqLambda"][i])
    return p


def n_ctus(width, height):
    return ((width + 63) // 64) * ((height + 63) // 64)


def oracle_compress(params, y, u, v, hints=None, wpp=False):
    """y,u,v: 2-D arrays (any integer dtype); hints: None or uint8 [numCtus][4] (0 none / 1 do not split / 2 split, the
    depth-1 CUs of a CTU in z-order).  Returns CTU_OUT_DT array."""
    n = n_ctus(params.width, params.height)
    out = np.zeros(n, CTU_OUT_DT)
    planes = [np.ascontiguousarray(p, dtype=np.uint16) for p in (y, u, v)]
    if wpp:
        if hints is not None:
            hints = np.ascontiguousarray(hints, dtype=np.uint8)
            assert hints.size == 4 * n
        rc = oracle().hm_intra_compress_picture_ex(C.byref(params), planes[0].ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data,
                                                   out.ctypes.data, hints.ctypes.data if hints is not None else None, 1)
    elif hints is None:
        rc = oracle().hm_intra_compress_picture(C.byref(params), planes[0].ctypes.data, planes[1].ctypes.data,
                                                planes[2].ctypes.data, out.ctypes.data)
    else:
        hints = np.ascontiguousarray(hints, dtype=np.uint8)
        assert hints.size == 4 * n
        rc = oracle().hm_intra_compress_picture_hinted(C.byref(params), planes[0].ctypes.data, planes[1].ctypes.data,
                                                       planes[2].ctypes.data, out.ctypes.data, hints.ctypes.data)
    assert rc == 0
    return out


def hints_from_split(width, height, split):
    """CNN decisions of one picture ([numCus] in the predictor's order: CTUs in raster order, the depth-1 CUs inside the
    picture in z-order) -> hint bytes [numCtus][4]: 1 = do not split, 2 = split, 0 for CUs cut by the picture edge."""
    wc, hc = (width + 63) // 64, (height + 63) // 64
    hints = np.zeros((wc * hc, 4), np.uint8)
    i = 0
    for cy in range(hc):
        for cx in range(wc):
            for k in range(4):
                x, y = cx * 64 + 32 * (k & 1), cy * 64 + 32 * (k >> 1)
                if x + 32 <= width and y + 32 <= height:
                    hints[cy * wc + cx, k] = 2 if split[i] else 1
                    i += 1
    assert i == len(split)
    return hints


def _oracle_job(job):
    (w, h, bd, qp), y, u, v = job[:4]
    return oracle_compress(params_from_qp(w, h, bd, qp), y, u, v, job[4] if len(job) > 4 else None, job[5] if len(job) > 5 else False)


def oracle_compress_parallel(jobs):
    """jobs: list of ((width, height, bitDepth, qp), y, u, v); one oracle process per picture (a 1080p picture takes
    ~20 s of one host core).  Returns the CTU_OUT_DT arrays in order."""
    import multiprocessing as mp
    build_oracle()
    n = min(len(jobs), max(1, len(os.sched_getaffinity(0))))
    if n <= 1:
        return [_oracle_job(j) for j in jobs]
    with mp.get_context("fork").Pool(n) as pool:
        return pool.map(_oracle_job, jobs)


def parse_dump(path):
    """Returns list of (slice_rec, ctu_recs) per picture."""
    raw = open(path, "rb").read()
    pos, pics = 0, []
    while pos < len(raw):
        magic = int(np.frombuffer(raw, "<u4", 1, pos)[0])
        if magic == 0x53434846:
            s = np.frombuffer(raw, SLICE_DT, 1, pos)[0]
            pos += SLICE_DT.itemsize
            pics.append((s, []))
        elif magic == 0x43434846:
            c = np.frombuffer(raw, CTU_DUMP_DT, 1, pos)[0]
            pos += CTU_DUMP_DT.itemsize
            pics[-1][1].append(c)
        else:
            raise ValueError("bad magic %x at %d" % (magic, pos))
    return [(s, np.array(c, CTU_DUMP_DT)) for s, c in pics]


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "TAppEncoderDump"))


def have_integrated():
    return os.path.exists(os.path.join(REF_DIR, "TAppEncoderFHC")) and os.path.exists(os.path.join(REF_DIR, "TAppEncoder"))


def build_emu_abi():
    """Test-only libfasthevc_cuda.so: the ABI subset the shim calls, on the host emulation build of the kernel source."""
    out_dir = os.path.join(ROOT, "tests", "_build", "abi")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libfasthevc_cuda.so")
    srcs = [os.path.join(ROOT, "tests", "emu_abi.cpp"), os.path.join(ROOT, "tests", "emu_harness.cpp"),
            os.path.join(ROOT, "fasthevc_b200", "csrc", "fhc_engine.cuh"), os.path.join(ROOT, "fasthevc_b200", "csrc", "fhc_rom.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, srcs[0], srcs[1], "-lm"])
    return out_dir


def run_reference(yuv_path, width, height, frames, qp, workdir, bit_depth=8, dump=True, extra=(), exe=None, lib_dir=None, tag="",
                  env_extra=None):
    """Runs oracle/_ref/TAppEncoder[Dump] (or `exe`); returns (bitstream bytes, dump path or None, stdout).
    lib_dir: directory put in front of the library search path (the integrated encoder's libfasthevc_cuda.so)."""
    exe = os.path.join(REF_DIR, exe or ("TAppEncoderDump" if dump else "TAppEncoder"))
    cfg = os.path.join(REF_DIR, "cfg", "encoder_intra_main.cfg" if bit_depth == 8 else "encoder_intra_main10.cfg")
    bs = os.path.join(workdir, "out%s.bin" % tag)
    dp = os.path.join(workdir, "dump.bin")
    env = dict(os.environ)
    if lib_dir:
        env["LD_LIBRARY_PATH"] = lib_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env.update(env_extra or {})
    if dump:
        env["FHC_DUMP"] = dp
    cmd = [exe, "-c", cfg, "-i", yuv_path, "-wdt", str(width), "-hgt", str(height), "-fr", "30", "-f", str(frames),
           "-q", str(qp), "-b", bs, "-o", os.path.join(workdir, "rec%s.yuv" % tag), "--SEIDecodedPictureHash=1"]
    if bit_depth != 8:
        cmd += ["--InputBitDepth=%d" % bit_depth]
    cmd += list(extra)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=workdir)
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-2000:] + r.stderr[-2000:])
    return open(bs, "rb").read(), (dp if dump else None), r.stdout


def compare_ctus(got, ref, what="", coef_from_dump=True):
    """Field-by-field comparison of CTU records; returns list of mismatch strings."""
    bad = []
    for i in range(len(ref)):
        for f in FIELDS:
            a, b = got[i][f], ref[i][f]
            if f == "totalCost":
                ok = float(a) == float(b)
            else:
                ok = np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))
            if not ok:
                bad.append("%s ctu %d field %s" % (what, i, f))
    return bad


# --- C-ABI structs (include/fasthevc_cuda.h) -------------------------------------------------
class FhcSliceParams(C.Structure):
    _fields_ = [("sliceQp", C.c_int32), ("reserved", C.c_int32), ("lambda_", C.c_double), ("sqrtLambda", C.c_double),
                ("chromaDistWeight", C.c_double * 2), ("rdoqLambda", C.c_double * 3)]


class FhcSeqParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "bitDepth", "chromaFormat", "ctuSize", "maxCuDepth", "tuLog2Max",
                                         "tuLog2Min", "tuMaxDepthIntra", "useRdoq", "useRdoqTs", "useTransformSkip",
                                         "useTransformSkipFast", "signHideFlag", "strongIntraSmoothing", "fastUdiUseMpm",
                                         "maxFramesInFlight", "waveFrontSynchro")]


def slice_params_from_hm(p):
    from fasthevc_b200.api import SliceParams
    sp = SliceParams()
    sp.sliceQp = p.qp
    sp.lambda_, sp.sqrtLambda = p.lambda_, p.sqrtLambda
    sp.chromaDistWeight[0], sp.chromaDistWeight[1] = p.distWeight[1], p.distWeight[2]
    for i in range(3):
        sp.rdoqLambda[i] = p.rdoqLambda[i]
    return sp


def build_emu():
    """Host emulation build of the kernel source (tests only)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "fhc_emu.so")
    srcs = [os.path.join(ROOT, "tests", "emu_harness.cpp"), os.path.join(ROOT, "fasthevc_b200", "csrc", "fhc_engine.cuh"),
            os.path.join(ROOT, "fasthevc_b200", "csrc", "fhc_rom.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, srcs[0], "-lm"])
    return so


_emu = None


def emu_compress(params, y, u, v, hints=None, wpp=False, order=0):
    global _emu
    if _emu is None:
        _emu = C.CDLL(build_emu())
        _emu.fhc_emu_compress_picture.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]
        _emu.fhc_emu_compress_picture_hinted.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _emu.fhc_emu_compress_picture_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    sb = 1 if params.bitDepth == 8 else 2
    dt = np.uint8 if sb == 1 else np.uint16
    planes = [np.ascontiguousarray(p, dtype=dt) for p in (y, u, v)]
    out = np.zeros(n_ctus(params.width, params.height), CTU_OUT_DT)
    sp = slice_params_from_hm(params)
    if wpp:
        if hints is not None:
            hints = np.ascontiguousarray(hints, dtype=np.uint8)
            assert hints.size == 4 * len(out)
        rc = _emu.fhc_emu_compress_picture_ex(params.width, params.height, params.bitDepth, sb, C.byref(sp), planes[0].ctypes.data,
                                              planes[1].ctypes.data, planes[2].ctypes.data, out.ctypes.data,
                                              hints.ctypes.data if hints is not None else None, 1, int(order))
    elif hints is None:
        rc = _emu.fhc_emu_compress_picture(params.width, params.height, params.bitDepth, sb, C.byref(sp), planes[0].ctypes.data,
                                           planes[1].ctypes.data, planes[2].ctypes.data, out.ctypes.data)
    else:
        hints = np.ascontiguousarray(hints, dtype=np.uint8)
        assert hints.size == 4 * len(out)
        rc = _emu.fhc_emu_compress_picture_hinted(params.width, params.height, params.bitDepth, sb, C.byref(sp), planes[0].ctypes.data,
                                                  planes[1].ctypes.data, planes[2].ctypes.data, out.ctypes.data, hints.ctypes.data)
    assert rc == 0
    return out


def decoder_md5_ok(bitstream_path, workdir):
    """TAppDecoder on the bitstream: number of pictures whose MD5 SEI matched, and whether any mismatched."""
    r = subprocess.run([os.path.join(REF_DIR, "TAppDecoder"), "-b", bitstream_path, "-o", os.path.join(workdir, "dec.yuv")],
                       capture_output=True, text=True, cwd=workdir)
    return r.stdout.count("(OK)"), ("ERROR" in r.stdout or r.returncode != 0)


# --- full-size expectations: per-CTU CRC-32 fixtures of the ORACLE's records ----------------------------------------
# A 2160p picture costs the oracle minutes of one host core; the GPU box's minutes are better spent on the GPU.  The
# oracle's records for the full-size test pictures are therefore computed in the container (tests/golden/
# make_fullsize_crc.py) and committed as one CRC-32 per CTU over every compared field; a mismatch falls back to the
# oracle itself to name the field.
FULLSIZE_CASES = {
    # tag: (width, height, bit depth, sigma, seed, [qp per picture], wavefront)
    "1080p": (1920, 1080, 8, 6.0, 1234, [22, 27, 32, 37], False),
    "2160p_main10": (3840, 2160, 10, 6.0, 31, [27], False),
    "1080p_wpp": (1920, 1080, 8, 6.0, 1234, [22, 27, 32, 37], True),
    "2160p_main10_wpp": (3840, 2160, 10, 6.0, 31, [27], True),
}


def ctu_crcs(recs):
    import zlib
    out = np.zeros(len(recs), np.uint32)
    for i in range(len(recs)):
        c = 0
        for f in FIELDS:
            a = recs[i][f]
            a = np.float64(a) if f == "totalCost" else np.asarray(a).astype(np.int32)
            c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
        out[i] = c
    return out


def fullsize_frames(tag):
    from fasthevc_b200.synth import synth_frames
    w, h, bd, sigma, seed, qps, wpp = FULLSIZE_CASES[tag]
    return synth_frames(w, h, len(qps), bit_depth=bd, sigma=sigma, seed=seed)


def fullsize_jobs(tag):
    w, h, bd, sigma, seed, qps, wpp = FULLSIZE_CASES[tag]
    Y, U, V = fullsize_frames(tag)
    return [((w, h, bd, qps[k]), Y[k], U[k], V[k], None, wpp) for k in range(len(qps))]


def check_fullsize(tag, got):
    """got: list of CTU record arrays, one per picture of the case.  Returns the list of mismatches."""
    path = os.path.join(GOLDEN, "fullsize_%s.npz" % tag)
    jobs = fullsize_jobs(tag)
    if os.path.exists(path):
        want = np.load(path)["crc"]
        assert want.shape == (len(got), len(got[0]))
        wrong = [k for k in range(len(got)) if not np.array_equal(ctu_crcs(got[k]), want[k])]
        if not wrong:
            return []
        bad = []
        for k in wrong:
            bad += compare_ctus(got[k], _oracle_job(jobs[k]), "%s pic%d" % (tag, k)) or ["%s pic%d: CRC differs" % (tag, k)]
        return bad
    ref = oracle_compress_parallel(jobs)
    bad = []
    for k in range(len(got)):
        bad += compare_ctus(got[k], ref[k], "%s pic%d" % (tag, k))
    return bad


# --- deblocking oracle (oracle/hm_deblock.c) --------------------------------------------------------------------------
_dbk = None


def oracle_deblock(width, height, bit_depth, qp, records):
    """TComLoopFilter::loopFilterPic restated on CTU records -> (Y, Cb, Cr) int16 planes."""
    global _dbk
    if _dbk is None:
        so = os.path.join(ORACLE_DIR, "libhm_deblock_oracle.so")
        src = os.path.join(ORACLE_DIR, "hm_deblock.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
        _dbk = C.CDLL(so)
        _dbk.hm_deblock_picture.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4
    recs = np.ascontiguousarray(records)
    assert recs.dtype == CTU_OUT_DT and len(recs) == n_ctus(width, height)
    y = np.zeros((height, width), np.int16)
    u = np.zeros((height // 2, width // 2), np.int16)
    v = np.zeros((height // 2, width // 2), np.int16)
    rc = _dbk.hm_deblock_picture(width, height, bit_depth, qp, recs.ctypes.data, y.ctypes.data, u.ctypes.data, v.ctypes.data)
    assert rc == 0
    return y, u, v


def planes_of(packed, width, height):
    """One picture in the planar 4:2:0 layout -> (Y, Cb, Cr) int16 planes."""
    p = np.asarray(packed).reshape(-1).astype(np.int16)
    l = width * height
    return p[:l].reshape(height, width), p[l:l + l // 4].reshape(height // 2, width // 2), p[l + l // 4:].reshape(height // 2, width // 2)


# --- SAO statistics oracle (oracle/hm_sao.c) --------------------------------------------------------------------------
_sao = None
SAO_HM_DT = np.dtype([("diff", "<i8", 32), ("count", "<i8", 32)])            # HM's SAOStatData
SAO_STATS_DT = np.dtype([("eoDiff", "<i4", (4, 5)), ("eoCount", "<i4", (4, 5)), ("boDiff", "<i4", 32), ("boCount", "<i4", 32)])   # fhc_sao_stats


def oracle_sao_stats(width, height, bit_depth, deblocked, original):
    """TEncSampleAdaptiveOffset::getStatistics restated; deblocked / original: (Y, Cb, Cr) planes.  Returns [numCtus][3][5] SAO_HM_DT."""
    global _sao
    if _sao is None:
        so = os.path.join(ORACLE_DIR, "libhm_sao_oracle.so")
        src = os.path.join(ORACLE_DIR, "hm_sao.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-o", so, src])
        _sao = C.CDLL(so)
        _sao.hm_sao_statistics.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3
    src_p = [np.ascontiguousarray(a, dtype=np.int16) for a in deblocked]
    org_p = [np.ascontiguousarray(a, dtype=np.int16) for a in original]
    out = np.zeros((n_ctus(width, height), 3, 5), SAO_HM_DT)
    sp = (C.c_void_p * 3)(*[a.ctypes.data for a in src_p])
    op = (C.c_void_p * 3)(*[a.ctypes.data for a in org_p])
    assert _sao.hm_sao_statistics(width, height, bit_depth, sp, op, out.ctypes.data) == 0
    return out


def sao_compact_to_hm(stats):
    """fhc_sao_stats records [numCtus][3] -> HM's layout [numCtus][3][5] (EO classes at diff[0..4], BO bands at diff[0..31])."""
    stats = np.asarray(stats)
    out = np.zeros(stats.shape + (5,), SAO_HM_DT)
    out["diff"][..., :4, :5] = stats["eoDiff"]
    out["count"][..., :4, :5] = stats["eoCount"]
    out["diff"][..., 4, :] = stats["boDiff"]
    out["count"][..., 4, :] = stats["boCount"]
    return out
