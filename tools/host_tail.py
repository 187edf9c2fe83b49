#!/usr/bin/env python3
"""Where the integrated encoder's time goes when the device is infinitely fast (build container, CPU only): runs
oracle/_ref/TAppEncoderFHC against tests/replay_abi.cpp, which serves precomputed CTU records (made here by the oracle)
without computing anything - what remains is the unchanged host's per-picture tail.
usage: python tools/host_tail.py [pictures] [qp]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fhc_testlib as T  # noqa: E402
from fasthevc_b200.synth import synth_frames, write_yuv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
qp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
w, h, distinct = 1920, 1080, 4
d = tempfile.mkdtemp()
out_dir = os.path.join(ROOT, "tests", "_build", "replay")
os.makedirs(out_dir, exist_ok=True)
subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-o", os.path.join(out_dir, "libfasthevc_cuda.so"),
                       os.path.join(ROOT, "tests", "replay_abi.cpp")])
Y, U, V = synth_frames(w, h, distinct, sigma=6.0)
idx = [k % distinct for k in range(n)]
write_yuv(os.path.join(d, "s.yuv"), Y[idx], U[idx], V[idx])
t0 = time.perf_counter()
recs = T.oracle_compress_parallel([((w, h, 8, qp), Y[k], U[k], V[k]) for k in range(distinct)])
print("oracle: %d pictures in %.1f s" % (distinct, time.perf_counter() - t0))
np.concatenate(recs).tofile(os.path.join(d, "records.bin"))
for tag, env in (("deblocking + SAO statistics of the kernel source, on the host", {}), ("the host's own deblocking and getStatistics (FHC_SHIM_DEBLOCK=0)", {"FHC_SHIM_DEBLOCK": "0"}),
                 ("deblocking on the host build, SAO statistics at no cost (zeros: timing only, another bitstream)", {"FHC_REPLAY_FAKE_SAO": "1"})):
    e = {"FHC_REPLAY": os.path.join(d, "records.bin"), "FHC_SHIM_INPUT": os.path.join(d, "s.yuv"), "FHC_SHIM_LOOKAHEAD": "4"}
    e.update(env)
    t0 = time.perf_counter()
    bs, _, log = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, dump=False, exe="TAppEncoderFHC", lib_dir=out_dir, env_extra=e, tag="_r")
    wall = time.perf_counter() - t0
    ok, bad = T.decoder_md5_ok(os.path.join(d, "out_r.bin"), d)
    print("%s: %d pictures 1080p QP %d in %.2f s = %.3f s per picture (%d bytes, decoder MD5 ok %d/%d)" % (tag, n, qp, wall, wall / n, len(bs), ok, n))
