"""Host pipeline: one device process, N host encoder processes, one bitstream.

The unchanged reference host is single-threaded and its per-picture tail (RD re-coding of the CTUs, SAO decision, final
CABAC, MD5: TEncGOP.cpp:1134-1870) costs about 0.12 s per 1080p picture even when the device delivers the mode decision,
the deblocked picture and the SAO statistics for free - so one host process caps the integrated encoder below what one
B200 delivers.  All-intra pictures are independent and, with the cfg's CRA pictures and ReWriteParamSetsFlag, every access
unit is self-contained, so the tail is spread over host PROCESSES:

  * this process owns the GPU: it reads the input, runs the mode decision of all pictures (pipelined submits), the
    deblocking and the SAO statistics (fhc_deblock_pictures_stats), and writes one spool file per picture;
  * N encoder processes (the reference encoder with the shim, oracle/_ref/TAppEncoderFHC, or what a maintainer builds per
    INTEGRATION.md), each on a contiguous frame group (--FrameSkip k --FramesToBeEncoded n, FHC_POC_BASE=k), load
    libfasthevc_spool.so in place of libfasthevc_cuda.so: it hands the device's results out of the spool and computes
    nothing; the processes start at once and wait for the pictures they need;
  * the groups' bitstreams are concatenated in order: byte-identical to the single-process bitstream
    (reference: the splice of TAppEncTop::xWriteOutput, TAppEncTop.cpp:629-693; tests/test_zpipeline.py).

Python here is host plumbing (process control, file I/O); the device work goes through the C ABI as everywhere else.
"""
import os
import shutil
import struct
import subprocess
import tempfile
import threading
import time

import numpy as np

from . import api
from .sharding import frame_group

_HERE = os.path.dirname(os.path.abspath(__file__))
SPOOL_MAGIC = 0x4C505346


def spool_library_dir():
    """Directory holding libfasthevc_cuda.so -> the spool library (what the encoder processes put on LD_LIBRARY_PATH)."""
    return os.path.join(_HERE, "spool")


def build_spool_library():
    """g++ -> fasthevc_b200/spool/libfasthevc_cuda.so (the file name the shim is linked against)."""
    d = spool_library_dir()
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "libfasthevc_cuda.so")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(_HERE, "csrc", "fasthevc_spool.cpp")])
    return so


def write_spool_picture(spool_dir, poc, width, height, bit_depth, wpp, slice_params, records, deblocked, sao_stats):
    """One picture's device results -> <spool_dir>/p<poc>.fhc (written under a temporary name, then renamed)."""
    sample_bytes = 1 if bit_depth == 8 else 2
    hdr = struct.pack("<8I", SPOOL_MAGIC, width, height, bit_depth, len(records), sample_bytes, poc, 1 if wpp else 0)
    tmp = os.path.join(spool_dir, ".p%08d.tmp" % poc)
    with open(tmp, "wb") as f:
        f.write(hdr)
        f.write(bytes(slice_params))
        f.write(np.ascontiguousarray(records).tobytes())
        f.write(np.ascontiguousarray(deblocked).tobytes())
        f.write(np.ascontiguousarray(sao_stats).tobytes())
    os.rename(tmp, os.path.join(spool_dir, "p%08d.fhc" % poc))


def produce(backend, read_frames, frames, qp, spool_dir, batch, wpp=False, first_frame=0):
    """The device side: pictures [first_frame, first_frame + frames) through `backend` (an api.Encoder or an object with
    its interface) in batches, two submits in flight; every finished picture goes to the spool.  read_frames(first, n)
    returns the packed pictures (n, frame_samples).  Returns the seconds spent."""
    t0 = time.perf_counter()
    sp = api.slice_params_from_qp(qp)
    w, h, bd = backend.seq.width, backend.seq.height, backend.seq.bitDepth
    groups = [(p, min(batch, first_frame + frames - p)) for p in range(first_frame, first_frame + frames, batch)]

    def submit(g):
        first, n = g
        yuv = np.ascontiguousarray(read_frames(first, n))
        backend.submit(yuv, [sp] * n, first_poc=first)
        return yuv                                        # kept alive until the group has been drained

    def drain(g):
        first, n = g
        planes, stats = backend.deblock_pictures_stats(first, n)      # waits for the group's pictures
        for k in range(n):
            write_spool_picture(spool_dir, first + k, w, h, bd, wpp, sp, backend.fetch_picture(first + k), planes[k], stats[k])

    if hasattr(backend, "set_pipelined") and len(groups) > 1 and not wpp:
        backend.set_pipelined(True)
    pending = submit(groups[0]) if groups else None
    for i, g in enumerate(groups):
        nxt = submit(groups[i + 1]) if i + 1 < len(groups) else None   # the next group is on the device before this one is drained
        drain(g)
        pending = nxt
    del pending
    return time.perf_counter() - t0


def encode(encoder_exe, cfg, yuv_path, width, height, frames, qp, out_path, workers, bit_depth=8, batch=None, wpp=False, device=0,
           backend=None, spool_dir=None, extra_args=(), keep_spool=False):
    """Encodes `frames` pictures of the planar 4:2:0 file `yuv_path` into the Annex-B file `out_path` with one device
    process (this one) and `workers` host encoder processes.  backend: None = api.Encoder on `device` (the CUDA library);
    tests pass an object with the same interface.  Returns a dict of timings."""
    sample_bytes = 1 if bit_depth == 8 else 2
    frame_bytes = width * height * 3 // 2 * sample_bytes
    own_spool = spool_dir is None
    if own_spool:
        spool_dir = tempfile.mkdtemp(prefix="fhc_spool_")
    lib_dir = spool_library_dir()
    if not os.path.exists(os.path.join(lib_dir, "libfasthevc_cuda.so")):
        build_spool_library()
    workers = max(1, min(workers, frames))
    if batch is None:
        batch = max(1, min(frames, api.pictures_in_flight(device) if backend is None else 8))
    own_backend = backend is None
    if own_backend:
        backend = api.Encoder(width, height, bit_depth, max_frames=batch, device=device, wpp=wpp)

    def read_frames(first, n):
        with open(yuv_path, "rb") as f:
            f.seek(first * frame_bytes)
            raw = f.read(n * frame_bytes)
        return np.frombuffer(raw, np.uint8 if sample_bytes == 1 else np.uint16).reshape(n, -1)

    # the host processes start at once: each waits in fhc_fetch_picture for the spool files of its frame group
    t0 = time.perf_counter()
    procs, parts = [], []
    for r in range(workers):
        first, last = frame_group(frames, r, workers)
        if last <= first:
            continue
        part = out_path + ".part%03d" % r
        parts.append(part)
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = lib_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        env["FHC_SPOOL_DIR"] = spool_dir
        env["FHC_POC_BASE"] = str(first)
        env.pop("FHC_SHIM_INPUT", None)                    # no read-ahead: nothing is submitted from there
        cmd = [encoder_exe, "-c", cfg, "-i", yuv_path, "-wdt", str(width), "-hgt", str(height), "-fr", "30", "-fs", str(first), "-f", str(last - first),
               "-q", str(qp), "-b", part, "-o", "", "--SEIDecodedPictureHash=1"]
        if bit_depth != 8:
            cmd += ["--InputBitDepth=%d" % bit_depth]
        if wpp:
            cmd += ["--WaveFrontSynchro=1"]
        cmd += list(extra_args)
        procs.append(subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    device_s = [0.0]
    err = []

    def device_side():
        try:
            device_s[0] = produce(backend, read_frames, frames, qp, spool_dir, batch, wpp)
        except Exception as e:                             # the host processes would wait for ever
            err.append(e)
            for p in procs:
                p.kill()

    t = threading.Thread(target=device_side)
    t.start()
    rcs = [(p.wait(), p.stderr.read().decode(errors="replace")[-500:]) for p in procs]
    t.join()
    if own_backend:
        backend.close()
    if err:
        raise err[0]
    for rc, tail in rcs:
        if rc != 0:
            raise api.FhcError("an encoder process failed (%d): %s" % (rc, tail))
    with open(out_path, "wb") as out:
        for part in parts:
            with open(part, "rb") as f:
                shutil.copyfileobj(f, out)
            os.remove(part)
    wall = time.perf_counter() - t0
    if own_spool and not keep_spool:
        shutil.rmtree(spool_dir, ignore_errors=True)
    return {"frames": frames, "workers": len(procs), "wall_s": wall, "device_s": device_s[0], "fps": frames / wall}
