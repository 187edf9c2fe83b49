#!/usr/bin/env python3
"""GPU measurement of fhc_deblock_pictures: 64 pictures 1920x1080 (searched with wavefronts, which is quick), deblocked
in one call; prints the CUDA-event time of the four kernels and the algorithmic GB/s."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames

W, H, N = 1920, 1080, 64
Y, U, V = synth_frames(W, H, 8)
idx = [k % 8 for k in range(N)]
enc = fh.Encoder(W, H, 8, max_frames=N, wpp=True)
yuv = enc.pinned_array((N, enc.frame_samples), np.uint8)
yuv[:] = enc.pack_frames(Y[idx], U[idx], V[idx])
enc.submit(yuv, [fh.slice_params_from_qp(37 if k % 2 else 32) for k in range(N)])
out = enc.pinned_array((N, enc.frame_samples), np.uint8)
for rep in range(4):
    _, ms = enc.deblock_pictures(0, N, out)
    algo = N * (enc.frame_samples * 3 + enc.num_ctus * 512)
    print("deblock %d x 1080p: kernels %.3f ms = %.0f pictures/s, %.0f GB/s algorithmic (records in + pictures out), checksum %d"
          % (N, ms, N / ms * 1e3, algo / ms / 1e6, int(out[::7, ::1001].sum(dtype=np.int64))), flush=True)
enc.close()
