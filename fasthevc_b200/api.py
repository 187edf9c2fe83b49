"""ctypes mirror of include/fasthevc_cuda.h (the TLibEncoderCUDA boundary)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class FhcError(RuntimeError):
    pass


class SliceParams(C.Structure):
    """fhc_slice_params: per-picture lambdas computed by the unchanged host (TEncSlice.cpp:112-144)."""
    _fields_ = [("sliceQp", C.c_int32), ("reserved", C.c_int32), ("lambda_", C.c_double), ("sqrtLambda", C.c_double),
                ("chromaDistWeight", C.c_double * 2), ("rdoqLambda", C.c_double * 3)]


class SeqParams(C.Structure):
    """fhc_seq_params."""
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "bitDepth", "chromaFormat", "ctuSize", "maxCuDepth", "tuLog2Max",
                                         "tuLog2Min", "tuMaxDepthIntra", "useRdoq", "useRdoqTs", "useTransformSkip",
                                         "useTransformSkipFast", "signHideFlag", "strongIntraSmoothing", "fastUdiUseMpm",
                                         "maxFramesInFlight", "waveFrontSynchro")]


# fhc_ctu_result
CTU_RESULT_DTYPE = np.dtype([("depth", "u1", 256), ("partSize", "u1", 256), ("predMode", "u1", 256), ("lumaDir", "u1", 256),
                             ("chromaDir", "u1", 256), ("trIdx", "u1", 256), ("cbf", "u1", (3, 256)), ("tskip", "u1", (3, 256)),
                             ("coefY", "<i2", 4096), ("coefCb", "<i2", 1024), ("coefCr", "<i2", 1024),
                             ("recoY", "<i2", 4096), ("recoCb", "<i2", 1024), ("recoCr", "<i2", 1024),
                             ("totalCost", "<f8"), ("totalDist", "<u4"), ("totalBits", "<u4"), ("totalBins", "<u4")], align=True)

# fhc_sao_stats
SAO_STATS_DTYPE = np.dtype([("eoDiff", "<i4", (4, 5)), ("eoCount", "<i4", (4, 5)), ("boDiff", "<i4", 32), ("boCount", "<i4", 32)])

EXPORTS = ["fhc_create", "fhc_destroy", "fhc_submit_frames", "fhc_fetch_ctu", "fhc_fetch_picture", "fhc_num_ctus",
           "fhc_slice_params_from_qp", "fhc_device_input", "fhc_run_resident", "fhc_host_alloc", "fhc_host_free",
           "fhc_last_error", "fhc_launch_count", "fhc_last_kernel_ms", "fhc_pictures_in_flight", "fhc_poll_picture",
           "fhc_fetch_picture_async", "fhc_fetch_sync", "fhc_measure_int_issue", "fhc_submit_frames_hinted", "fhc_device_picture", "fhc_configure", "fhc_span_begin", "fhc_span_ms", "fhc_deblock_pictures", "fhc_last_deblock_ms", "fhc_deblock_pictures_stats"]


def library_path():
    return os.environ.get("FHC_LIBRARY", os.path.join(_HERE, "libfasthevc_cuda.so"))


def build_library(force=False, verbose=False):
    """nvcc cross-compiles for sm_100a; the .so stays in-tree (git-ignored) and travels with the snapshot."""
    so = library_path()
    srcs = [os.path.join(_CSRC, f) for f in ("fasthevc_cuda.cu", "fhc_engine.cuh", "fhc_rom.h")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "fasthevc_cuda.h"))
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", so, srcs[0]]
    subprocess.check_call(cmd)
    return so


_lib = None


def load_library():
    """Loads libfasthevc_cuda.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    so = library_path()
    if not os.path.exists(so):
        raise FhcError("libfasthevc_cuda.so is missing: run __graft_entry__.build() (nvcc) first; there is no CPU path")
    lib = C.CDLL(so)
    lib.fhc_create.argtypes = [C.POINTER(SeqParams), C.c_int, C.POINTER(C.c_void_p)]
    lib.fhc_destroy.argtypes = [C.c_void_p]
    lib.fhc_submit_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(SliceParams)]
    lib.fhc_submit_frames_hinted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(SliceParams), C.c_void_p]
    lib.fhc_deblock_pictures.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.fhc_deblock_pictures_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_last_deblock_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.fhc_configure.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.fhc_span_begin.argtypes = [C.c_void_p]
    lib.fhc_span_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.fhc_device_picture.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.fhc_fetch_ctu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.fhc_fetch_picture.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_num_ctus.argtypes = [C.c_void_p]
    lib.fhc_slice_params_from_qp.argtypes = [C.c_int, C.POINTER(SliceParams)]
    lib.fhc_device_input.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.fhc_run_resident.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(SliceParams), C.POINTER(C.c_float)]
    lib.fhc_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.fhc_host_free.argtypes = [C.c_void_p]
    lib.fhc_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.fhc_pictures_in_flight.argtypes = [C.c_int]
    lib.fhc_poll_picture.argtypes = [C.c_void_p, C.c_int]
    lib.fhc_fetch_picture_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.fhc_fetch_sync.argtypes = [C.c_void_p]
    lib.fhc_measure_int_issue.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.fhc_last_error.restype = C.c_char_p
    lib.fhc_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def pictures_in_flight(device=0):
    """Pictures the persistent kernel keeps in flight on `device` (a submit of that many fills the GPU)."""
    n = load_library().fhc_pictures_in_flight(int(device))
    if n <= 0:
        raise FhcError("fhc_pictures_in_flight(%d) -> %d: %s" % (device, n, (load_library().fhc_last_error() or b"").decode()))
    return n


def measure_int_issue(device=0):
    """(Gop/s, ms): measured int32 issue roof of the device (IMAD / IADD3 / LOP3 mix at full occupancy)."""
    g, ms = C.c_double(), C.c_double()
    rc = load_library().fhc_measure_int_issue(int(device), C.byref(g), C.byref(ms))
    if rc != 0:
        raise FhcError("fhc_measure_int_issue -> %d: %s" % (rc, (load_library().fhc_last_error() or b"").decode()))
    return g.value, ms.value


def slice_params_from_qp(qp):
    sp = SliceParams()
    rc = load_library().fhc_slice_params_from_qp(int(qp), C.byref(sp))
    if rc != 0:
        raise FhcError("fhc_slice_params_from_qp(%d) -> %d" % (qp, rc))
    return sp


class Encoder:
    """One fhc_ctx: the mode-decision engine for one sequence geometry on one device."""

    def __init__(self, width, height, bit_depth=8, max_frames=1, device=0, wpp=False):
        """wpp: --WaveFrontSynchro=1 (the CTU rows of a picture are searched concurrently; results = the reference's with that flag)."""
        self.lib = load_library()
        sq = SeqParams(width, height, bit_depth, 1, 64, 4, 5, 2, 3, 1, 1, 1, 1, 1, 1, 1, max_frames, 1 if wpp else 0)
        self.seq = sq
        self.ctx = C.c_void_p()
        self._check(self.lib.fhc_create(C.byref(sq), device, C.byref(self.ctx)), "fhc_create")
        self.num_ctus = self.lib.fhc_num_ctus(self.ctx)
        self.sample_bytes = 1 if bit_depth == 8 else 2
        self.frame_samples = width * height * 3 // 2
        self._pinned = []

    def _check(self, rc, what):
        if rc != 0:
            raise FhcError("%s failed (%d): %s" % (what, rc, (self.lib.fhc_last_error() or b"").decode()))

    def close(self):
        if self.ctx:
            for p in self._pinned:
                self.lib.fhc_host_free(p)
            self._pinned = []
            self.lib.fhc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pinned_array(self, shape, dtype):
        """numpy view on page-locked host memory (cudaHostAlloc)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self.lib.fhc_host_alloc(C.byref(p), n), "fhc_host_alloc")
        self._pinned.append(p)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    @staticmethod
    def pack_frames(Y, U, V):
        """(n,H,W),(n,H/2,W/2)x2 -> one planar 4:2:0 buffer, picture after picture."""
        n = Y.shape[0]
        return np.concatenate([Y.reshape(n, -1), U.reshape(n, -1), V.reshape(n, -1)], axis=1)

    def submit(self, yuv, slice_params, first_poc=0, hints=None):
        """yuv: contiguous (n, frame_samples) uint8/uint16 host array; slice_params: list of SliceParams; hints: None or
        uint8 [n][numCtus][4] CU-split hints (fhc_submit_frames_hinted: 0 none, 1 do not split, 2 split)."""
        n = len(slice_params)
        arr = (SliceParams * n)(*slice_params)
        assert yuv.flags["C_CONTIGUOUS"] and yuv.dtype.itemsize == self.sample_bytes and yuv.size == n * self.frame_samples
        if hints is None:
            self._check(self.lib.fhc_submit_frames(self.ctx, first_poc, n, yuv.ctypes.data, self.sample_bytes, arr), "fhc_submit_frames")
            return
        hints = np.ascontiguousarray(hints, dtype=np.uint8)
        if hints.size != n * self.num_ctus * 4:
            raise FhcError("hints: %d bytes, expected %d" % (hints.size, n * self.num_ctus * 4))
        self._check(self.lib.fhc_submit_frames_hinted(self.ctx, first_poc, n, yuv.ctypes.data, self.sample_bytes, arr, hints.ctypes.data),
                    "fhc_submit_frames_hinted")

    def deblock_pictures_stats(self, first_poc, n, out=None):
        """fhc_deblock_pictures_stats: (deblocked pictures, SAO statistics [n][numCtus][3] of SAO_STATS_DTYPE)."""
        if out is None:
            out = np.zeros((n, self.frame_samples), np.uint8 if self.sample_bytes == 1 else np.uint16)
        stats = np.zeros((n, self.num_ctus, 3), SAO_STATS_DTYPE)
        self._check(self.lib.fhc_deblock_pictures_stats(self.ctx, int(first_poc), int(n), out.ctypes.data, self.sample_bytes, stats.ctypes.data),
                    "fhc_deblock_pictures_stats")
        return out, stats

    def deblock_pictures(self, first_poc, n, out=None):
        """fhc_deblock_pictures: the deblocked pictures [first_poc, first_poc + n) as an (n, frame_samples) array in the
        input layout; also returns the CUDA-event time of the kernels."""
        if out is None:
            out = np.zeros((n, self.frame_samples), np.uint8 if self.sample_bytes == 1 else np.uint16)
        self._check(self.lib.fhc_deblock_pictures(self.ctx, int(first_poc), int(n), out.ctypes.data, self.sample_bytes), "fhc_deblock_pictures")
        ms = C.c_float()
        self._check(self.lib.fhc_last_deblock_ms(self.ctx, C.byref(ms)), "fhc_last_deblock_ms")
        return out, ms.value

    def set_pipelined(self, on=True):
        """FHC_OPT_PIPELINED_SUBMITS: the caller keeps a second submit queued behind the running one."""
        self._check(self.lib.fhc_configure(self.ctx, 1, 1 if on else 0), "fhc_configure")

    def span_begin(self):
        self._check(self.lib.fhc_span_begin(self.ctx), "fhc_span_begin")

    def span_ms(self):
        ms = C.c_float()
        self._check(self.lib.fhc_span_ms(self.ctx, C.byref(ms)), "fhc_span_ms")
        return ms.value

    def device_picture_ptr(self, poc):
        """Device address of submitted picture `poc` (and the bytes from there to the end of its submit)."""
        p, b = C.c_void_p(), C.c_uint64()
        self._check(self.lib.fhc_device_picture(self.ctx, int(poc), C.byref(p), C.byref(b)), "fhc_device_picture")
        return p.value, b.value

    def fetch_picture(self, poc, out=None):
        if out is None:
            out = np.zeros(self.num_ctus, CTU_RESULT_DTYPE)
        self._check(self.lib.fhc_fetch_picture(self.ctx, poc, out.ctypes.data), "fhc_fetch_picture")
        return out

    def poll_picture(self, poc):
        """True when picture `poc` is finished on the device (non-blocking)."""
        rc = self.lib.fhc_poll_picture(self.ctx, poc)
        if rc < 0:
            self._check(rc, "fhc_poll_picture")
        return rc == 1

    def fetch_picture_async(self, poc, out):
        """Waits for picture `poc` only, then queues its device->host copy into `out` (page-locked); fetch_sync() completes it."""
        self._check(self.lib.fhc_fetch_picture_async(self.ctx, poc, out.ctypes.data), "fhc_fetch_picture_async")

    def fetch_sync(self):
        self._check(self.lib.fhc_fetch_sync(self.ctx), "fhc_fetch_sync")

    def fetch_ctu(self, poc, addr):
        out = np.zeros(1, CTU_RESULT_DTYPE)
        self._check(self.lib.fhc_fetch_ctu(self.ctx, poc, addr, out.ctypes.data), "fhc_fetch_ctu")
        return out[0]

    def device_input_ptr(self):
        p, b = C.c_void_p(), C.c_uint64()
        self._check(self.lib.fhc_device_input(self.ctx, C.byref(p), C.byref(b)), "fhc_device_input")
        return p.value, b.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self.lib.fhc_last_kernel_ms(self.ctx, C.byref(ms)), "fhc_last_kernel_ms")
        return ms.value

    def run_resident(self, slice_params):
        n = len(slice_params)
        arr = (SliceParams * n)(*slice_params)
        ms = C.c_float()
        self._check(self.lib.fhc_run_resident(self.ctx, n, self.sample_bytes, arr, C.byref(ms)), "fhc_run_resident")
        return ms.value
