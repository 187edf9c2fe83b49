"""Host pipeline (fasthevc_b200/pipeline.py): one device process + N host encoder processes on
contiguous frame groups, the device's results handed over through spool files (libfasthevc_spool.so computes nothing),
the groups' bitstreams concatenated - byte-identical to the stock reference encoder's single-process bitstream.
  CPU: the device side played by the host build of the kernel source (tests/emu_abi.cpp) behind the Encoder interface.
  GPU: the same flow with the CUDA library.
"""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200 import api, pipeline
from fasthevc_b200.synth import synth_frames, write_yuv

pytestmark = pytest.mark.skipif(not T.have_integrated(), reason="oracle/_ref/TAppEncoderFHC not built")


class EmuBackend:
    """tests/emu_abi.cpp behind the interface of api.Encoder that pipeline.produce uses."""

    class Seq:
        pass

    def __init__(self, width, height, bit_depth, max_frames, wpp=False):
        self.lib = C.CDLL(os.path.join(T.build_emu_abi(), "libfasthevc_cuda.so"))
        self.lib.fhc_create.argtypes = [C.POINTER(T.FhcSeqParams), C.c_int, C.POINTER(C.c_void_p)]
        self.lib.fhc_submit_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.fhc_fetch_picture.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.lib.fhc_deblock_pictures_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.fhc_destroy.argtypes = [C.c_void_p]
        self.seq = self.Seq()
        self.seq.width, self.seq.height, self.seq.bitDepth = width, height, bit_depth
        sq = T.FhcSeqParams(width, height, bit_depth, 1, 64, 4, 5, 2, 3, 1, 1, 1, 1, 1, 1, 1, max_frames, 1 if wpp else 0)
        self.ctx = C.c_void_p()
        assert self.lib.fhc_create(C.byref(sq), 0, C.byref(self.ctx)) == 0
        self.num_ctus = T.n_ctus(width, height)
        self.sample_bytes = 1 if bit_depth == 8 else 2
        self.frame_samples = width * height * 3 // 2
        self.submits = []

    def submit(self, yuv, slice_params, first_poc=0):
        n = len(slice_params)
        arr = (api.SliceParams * n)(*slice_params)
        self.submits.append((first_poc, n))
        assert self.lib.fhc_submit_frames(self.ctx, first_poc, n, yuv.ctypes.data, self.sample_bytes, arr) == 0

    def fetch_picture(self, poc):
        out = np.zeros(self.num_ctus, api.CTU_RESULT_DTYPE)
        assert self.lib.fhc_fetch_picture(self.ctx, poc, out.ctypes.data) == 0
        return out

    def deblock_pictures_stats(self, first, n):
        out = np.zeros((n, self.frame_samples), np.uint8 if self.sample_bytes == 1 else np.uint16)
        stats = np.zeros((n, self.num_ctus, 3), api.SAO_STATS_DTYPE)
        assert self.lib.fhc_deblock_pictures_stats(self.ctx, first, n, out.ctypes.data, self.sample_bytes, stats.ctypes.data) == 0
        return out, stats

    def close(self):
        self.lib.fhc_destroy(self.ctx)


def _run(w, h, n, qp, bd, workers, batch, wpp, backend_factory, sigma=9.0):
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=w + n)
    yuv = os.path.join(d, "s.yuv")
    write_yuv(yuv, Y, U, V)
    extra = ["--WaveFrontSynchro=1"] if wpp else []
    ref, _, _ = T.run_reference(yuv, w, h, n, qp, d, bit_depth=bd, dump=False, tag="_ref", extra=extra)
    backend = backend_factory(w, h, bd, batch, wpp)
    try:
        info = pipeline.encode(os.path.join(T.REF_DIR, "TAppEncoderFHC"),
                               os.path.join(T.REF_DIR, "cfg", "encoder_intra_main.cfg" if bd == 8 else "encoder_intra_main10.cfg"),
                               yuv, w, h, n, qp, os.path.join(d, "pipe.bin"), workers, bit_depth=bd, batch=batch, wpp=wpp, backend=backend)
    finally:
        backend.close()
    got = open(os.path.join(d, "pipe.bin"), "rb").read()
    assert got == ref
    ok, bad = T.decoder_md5_ok(os.path.join(d, "pipe.bin"), d)
    assert ok == n and not bad
    assert not [f for f in os.listdir(d) if ".part" in f]
    return info, backend


@pytest.mark.parametrize("w,h,n,qp,bd,workers,batch,wpp", [(136, 72, 7, 31, 8, 3, 3, False), (136, 72, 5, 28, 10, 2, 2, False), (200, 136, 4, 33, 8, 4, 4, True)])
def test_pipeline_bitstream_is_byte_identical_emulated(w, h, n, qp, bd, workers, batch, wpp):
    pipeline.build_spool_library()
    info, backend = _run(w, h, n, qp, bd, workers, batch, wpp, lambda *a: EmuBackend(*a))
    assert info["workers"] == min(workers, n) and info["frames"] == n
    # batches went to the device in order, each exactly once
    assert backend.submits == [(p, min(batch, n - p)) for p in range(0, n, batch)]


def test_spool_library_rejects_foreign_results():
    """The spool library hands out only what matches the sequence and the lambdas this host computed: a picture searched with
    another QP makes the encoder process fail instead of writing a wrong bitstream."""
    pipeline.build_spool_library()
    w, h, n, qp = 136, 72, 2, 30
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, n, sigma=9.0, seed=4)
    yuv = os.path.join(d, "s.yuv")
    write_yuv(yuv, Y, U, V)
    backend = EmuBackend(w, h, 8, n)
    spool = os.path.join(d, "spool")
    os.makedirs(spool)
    raw = np.fromfile(yuv, np.uint8).reshape(n, -1)
    pipeline.produce(backend, lambda first, k: raw[first:first + k], n, qp + 3, spool, n)      # the device side used QP 33
    backend.close()
    with pytest.raises(RuntimeError):
        T.run_reference(yuv, w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=pipeline.spool_library_dir(),
                        env_extra={"FHC_SPOOL_DIR": spool, "FHC_SPOOL_TIMEOUT_S": "5"})
    with pytest.raises(RuntimeError):       # and a picture that never arrives is a time-out, not a hang
        T.run_reference(yuv, w, h, n + 1, qp + 3, d, dump=False, exe="TAppEncoderFHC", lib_dir=pipeline.spool_library_dir(),
                        env_extra={"FHC_SPOOL_DIR": os.path.join(d, "empty"), "FHC_SPOOL_TIMEOUT_S": "1"})


@pytest.mark.gpu
@pytest.mark.parametrize("wpp", [False, True], ids=["plain", "wavefront"])
def test_pipeline_bitstream_is_byte_identical_cuda(wpp):
    """The same flow with the CUDA library as the device side: 416x240, 10 pictures in batches of 4, 3 host processes."""
    import fasthevc_b200 as fh
    pipeline.build_spool_library()
    info, _ = _run(416, 240, 10, 32, 8, 3, 4, wpp, lambda w, h, bd, batch, wp: fh.Encoder(w, h, bd, max_frames=batch, device=0, wpp=wp), sigma=8.0)
    print("pipeline %s: %d pictures, %d host processes, %.2f s (device side %.2f s)" % ("wavefront" if wpp else "plain", info["frames"], info["workers"],
                                                                                        info["wall_s"], info["device_s"]))
