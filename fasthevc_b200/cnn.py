"""ctypes mirror of include/fasthevc_cnn.h: the CNN CU-split predictor.

The product is libfasthevc_cnn.so (fasthevc_b200/csrc/fhc_cnn.cu: tcgen05 convolutions).  No CPU path: every call raises
when the library or an sm_100 device is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .api import FhcError

_HERE = os.path.dirname(os.path.abspath(__file__))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = ["fhc_cnn_create", "fhc_cnn_destroy", "fhc_cnn_num_cus", "fhc_cnn_cu_positions", "fhc_cnn_predict",
           "fhc_cnn_predict_device", "fhc_cnn_last_kernel_ms", "fhc_cnn_rerun_resident", "fhc_cnn_last_error",
           "fhc_cnn_launch_count"]

_PF = C.POINTER(C.c_float)


class CnnWeights(C.Structure):
    """fhc_cnn_weights."""
    _fields_ = [("mean", _PF), ("conv_w", _PF * 3), ("conv_b", _PF * 3), ("bn_gamma", _PF * 3), ("bn_beta", _PF * 3),
                ("bn_mean", _PF * 3), ("bn_var", _PF * 3), ("bn_eps", C.c_float), ("reserved", C.c_float), ("fc_w", _PF),
                ("fc_b", _PF)]


def library_path():
    return os.environ.get("FHC_CNN_LIBRARY", os.path.join(_HERE, "libfasthevc_cnn.so"))


def build_library(verbose=False):
    """nvcc cross-compiles for sm_100a; always compiles (4 s)."""
    so = library_path()
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", so, os.path.join(_HERE, "csrc", "fhc_cnn.cu")]
    subprocess.check_call(cmd)
    return so


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    so = library_path()
    if not os.path.exists(so):
        raise FhcError("libfasthevc_cnn.so is missing: run __graft_entry__.build() (nvcc) first; there is no CPU path")
    lib = C.CDLL(so)
    lib.fhc_cnn_create.argtypes = [C.POINTER(CnnWeights), C.c_int, C.POINTER(C.c_void_p)]
    lib.fhc_cnn_destroy.argtypes = [C.c_void_p]
    lib.fhc_cnn_num_cus.argtypes = [C.c_int, C.c_int]
    lib.fhc_cnn_cu_positions.argtypes = [C.c_int, C.c_int, C.c_void_p]
    for f in (lib.fhc_cnn_predict, lib.fhc_cnn_predict_device):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.fhc_cnn_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    lib.fhc_cnn_rerun_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    lib.fhc_cnn_debug_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.fhc_cnn_last_error.restype = C.c_char_p
    lib.fhc_cnn_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _err(what, rc):
    return FhcError("%s -> %d: %s" % (what, rc, (load_library().fhc_cnn_last_error() or b"").decode()))


def num_cus(width, height):
    return load_library().fhc_cnn_num_cus(int(width), int(height))


def cu_positions(width, height):
    """[numCus][2] int32 (x, y) of the depth-1 CUs inside the picture, in the interface's order."""
    n = num_cus(width, height)
    xy = np.zeros((max(n, 0), 2), np.int32)
    if n > 0:
        got = load_library().fhc_cnn_cu_positions(int(width), int(height), xy.ctypes.data)
        assert got == n
    return xy


class SplitPredictor:
    """One predictor instance = weights folded and resident on one device.

    `weights`: dict in the layouts PyTorch / MATLAB export:
    mean [32][32]; conv{i}_w [Cout][Cin][3][3], conv{i}_b, bn{i}_gamma/beta/mean/var [Cout] for i = 0..2; bn_eps;
    fc_w [2][4096] (c, y, x), fc_b [2]."""

    def __init__(self, weights, device=0):
        lib = load_library()
        self._keep = []

        def ptr(a, n):
            a = np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))
            if a.size != n:
                raise FhcError("weight array has %d elements, expected %d" % (a.size, n))
            self._keep.append(a)
            return a.ctypes.data_as(_PF)

        w = CnnWeights()
        w.mean = ptr(weights["mean"], 1024)
        for i, (cin, cout) in enumerate(((3, 16), (16, 32), (32, 64))):
            w.conv_w[i] = ptr(weights["conv%d_w" % i], cout * cin * 9)
            w.conv_b[i] = ptr(weights["conv%d_b" % i], cout)
            w.bn_gamma[i] = ptr(weights["bn%d_gamma" % i], cout)
            w.bn_beta[i] = ptr(weights["bn%d_beta" % i], cout)
            w.bn_mean[i] = ptr(weights["bn%d_mean" % i], cout)
            w.bn_var[i] = ptr(weights["bn%d_var" % i], cout)
        w.bn_eps = float(weights["bn_eps"])
        w.fc_w = ptr(weights["fc_w"], 2 * 4096)
        w.fc_b = ptr(weights["fc_b"], 2)
        h = C.c_void_p()
        rc = lib.fhc_cnn_create(C.byref(w), int(device), C.byref(h))
        if rc != 0:
            if h:
                lib.fhc_cnn_destroy(h)
            raise _err("fhc_cnn_create", rc)
        self._h = h
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fhc_cnn_destroy(self._h)
            self._h = None

    __del__ = close

    def predict(self, yuv, sample_bytes, bit_depth, width, height, n):
        """yuv: host buffer in the layout of Encoder.pack_frames (numpy array or address).  Returns (logits [n][cus][2]
        fp32, split [n][cus] uint8)."""
        cus = num_cus(width, height)
        logits = np.zeros((n, cus, 2), np.float32)
        split = np.zeros((n, cus), np.uint8)
        addr = yuv.ctypes.data if isinstance(yuv, np.ndarray) else int(yuv)
        rc = self._lib.fhc_cnn_predict(self._h, addr, int(sample_bytes), int(bit_depth), int(width), int(height), int(n),
                                       logits.ctypes.data, split.ctypes.data)
        if rc != 0:
            raise _err("fhc_cnn_predict", rc)
        return logits, split

    def predict_device(self, dptr, sample_bytes, bit_depth, width, height, n):
        cus = num_cus(width, height)
        logits = np.zeros((n, cus, 2), np.float32)
        split = np.zeros((n, cus), np.uint8)
        rc = self._lib.fhc_cnn_predict_device(self._h, int(dptr), int(sample_bytes), int(bit_depth), int(width), int(height),
                                              int(n), logits.ctypes.data, split.ctypes.data)
        if rc != 0:
            raise _err("fhc_cnn_predict_device", rc)
        return logits, split

    def last_kernel_ms(self):
        ms, cus = C.c_float(), C.c_int64()
        rc = self._lib.fhc_cnn_last_kernel_ms(self._h, C.byref(ms), C.byref(cus))
        if rc != 0:
            raise _err("fhc_cnn_last_kernel_ms", rc)
        return ms.value, cus.value

    def rerun_resident(self, reps):
        ms = C.c_float()
        rc = self._lib.fhc_cnn_rerun_resident(self._h, int(reps), C.byref(ms))
        if rc != 0:
            raise _err("fhc_cnn_rerun_resident", rc)
        return ms.value

    def debug_dump(self):
        """FHC_CNN_DEBUG=1 at construction: (act1 [2][432][8] f16, act2 [4][160][8] f16, feat [64][64] f32) of CU 0."""
        nbytes = 2 * 432 * 16 + 4 * 160 * 16 + 4096 * 4
        buf = np.zeros(nbytes, np.uint8)
        rc = self._lib.fhc_cnn_debug_dump(self._h, buf.ctypes.data, nbytes)
        if rc < 0:
            raise _err("fhc_cnn_debug_dump", rc)
        a1 = buf[:2 * 432 * 16].view(np.float16).reshape(2, 432, 8)
        a2 = buf[2 * 432 * 16:2 * 432 * 16 + 4 * 160 * 16].view(np.float16).reshape(4, 160, 8)
        ft = buf[2 * 432 * 16 + 4 * 160 * 16:].view(np.float32).reshape(64, 64)
        return a1, a2, ft


def hints_from_split(width, height, split):
    """CNN decisions [n][numCus] (the predictor's order: CTUs in raster order, the depth-1 CUs inside the picture in
    z-order) -> hint bytes [n][numCtus][4] for fhc_submit_frames_hinted: 1 = do not split, 2 = split, 0 = CU cut by the
    picture edge (HM splits those without a decision, TEncCu.cpp:519-521)."""
    split = np.asarray(split, np.uint8)
    if split.ndim == 1:
        split = split[None]
    wc, hc = (width + 63) // 64, (height + 63) // 64
    xy = cu_positions(width, height)
    if split.shape[1] != len(xy):
        raise FhcError("split has %d entries per picture, expected %d" % (split.shape[1], len(xy)))
    hints = np.zeros((split.shape[0], wc * hc, 4), np.uint8)
    ctu = (xy[:, 1] // 64) * wc + xy[:, 0] // 64
    k = (xy[:, 0] % 64) // 32 + 2 * ((xy[:, 1] % 64) // 32)
    hints[:, ctu, k] = 1 + split
    return hints


def seeded_weights(seed=20240):
    """Random-init weights of the architecture (He-initialised convolutions, identity batch normalisation): for benchmarks
    and smoke runs - nothing trained exists.  The activations are kept O(1) by scaling each layer's BN
    variance to the fan-in; the two logits are NOT centred: see bench.py, which centres them on its own input."""
    rng = np.random.default_rng(seed)
    w = {"bn_eps": np.float32(1e-5), "mean": np.full((32, 32), 128.0, np.float32)}
    scale = 40.0                                    # typical deviation of a luma sample from the mean image
    for i, (cin, cout) in enumerate(((3, 16), (16, 32), (32, 64))):
        w["conv%d_w" % i] = (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(np.float32)
        w["conv%d_b" % i] = np.zeros(cout, np.float32)
        w["bn%d_gamma" % i] = np.ones(cout, np.float32)
        w["bn%d_beta" % i] = np.zeros(cout, np.float32)
        w["bn%d_mean" % i] = np.zeros(cout, np.float32)
        w["bn%d_var" % i] = np.full(cout, scale * scale if i == 0 else 1.0, np.float32)
    w["fc_w"] = (rng.standard_normal((2, 4096)) / 64.0).astype(np.float32)
    w["fc_b"] = np.zeros(2, np.float32)
    return w
