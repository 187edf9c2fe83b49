#!/usr/bin/env python3
"""GPU measurement of the wavefront variant (--WaveFrontSynchro=1): the bench workload (256 1080p jobs), one 1080p picture
alone, and BASELINE configs[3] (32 pictures 3840x2160 Main10, QP 27).  usage: python tools/wpp_check.py [1080 | 2160 ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames

what = sys.argv[1:] or ["1080", "2160"]


def run(enc, host_in, sps, poc, tag):
    t0 = time.perf_counter()
    enc.submit(host_in, sps, first_poc=poc)
    out = np.zeros(enc.num_ctus, fh.CTU_RESULT_DTYPE)
    bits, first = 0, None
    for j in range(len(sps)):
        enc.fetch_picture(poc + j, out)
        if first is None:
            first = time.perf_counter() - t0
        bits += int(out["totalBits"].sum(dtype=np.int64))
    ms = enc.last_kernel_ms()
    print("%s: %d pictures, kernel %.1f ms = %.2f fps, %.0f CTU/s, first picture after %.2f s, total bits %d"
          % (tag, len(sps), ms, len(sps) / ms * 1e3, len(sps) * enc.num_ctus / ms * 1e3, first, bits), flush=True)


if "1080" in what:
    W, H, N = 1920, 1080, 64
    Y, U, V = synth_frames(W, H, N)
    jobs = [(f, q) for q in (22, 27, 32, 37) for f in range(N)]
    enc = fh.Encoder(W, H, 8, max_frames=len(jobs), wpp=True)
    host_in = enc.pinned_array((len(jobs), enc.frame_samples), np.uint8)
    packed = enc.pack_frames(Y, U, V)
    for j, (f, _) in enumerate(jobs):
        host_in[j] = packed[f]
    sps = [fh.slice_params_from_qp(q) for _, q in jobs]
    run(enc, host_in, sps, 0, "1080p wavefront, bench workload")
    run(enc, host_in, sps, 1000, "1080p wavefront, bench workload (again)")
    for q, j in ((22, 0), (32, 128)):
        run(enc, host_in[j:j + 1], sps[j:j + 1], 2000 + j, "1080p wavefront, one QP-%d picture alone" % q)
    run(enc, host_in[:16], sps[:16], 3000, "1080p wavefront, 16 QP-22 pictures")
    enc.close()
if "2160" in what:
    W, H, N = 3840, 2160, 32
    Y, U, V = synth_frames(W, H, 4, bit_depth=10)
    enc = fh.Encoder(W, H, 10, max_frames=N, wpp=True)
    host_in = enc.pinned_array((N, enc.frame_samples), np.uint16)
    packed = enc.pack_frames(Y, U, V)
    for j in range(N):
        host_in[j] = packed[j % 4]
    sps = [fh.slice_params_from_qp(27)] * N
    run(enc, host_in, sps, 0, "2160p Main10 wavefront, configs[3]")
    run(enc, host_in[:1], sps[:1], 100, "2160p Main10 wavefront, one picture alone")
    enc.close()
