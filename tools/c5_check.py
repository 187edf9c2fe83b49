#!/usr/bin/env python3
"""GPU experiment: BASELINE configs[4] (256 1080p jobs with CNN hints) with one or two picture chains per block
(FHC_HINTED_CHAINS).  usage: python tools/c5_check.py [chains ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasthevc_b200 as fh
from fasthevc_b200 import cnn
from fasthevc_b200.synth import synth_frames

W, H, N = 1920, 1080, 64
QPS = [22, 27, 32, 37]
Y, U, V = synth_frames(W, H, N)
jobs = [(f, q) for q in QPS for f in range(N)]
enc = fh.Encoder(W, H, 8, max_frames=len(jobs))
host_in = enc.pinned_array((len(jobs), enc.frame_samples), np.uint8)
packed = enc.pack_frames(Y, U, V)
for j, (f, _) in enumerate(jobs):
    host_in[j] = packed[f]
sps = [fh.slice_params_from_qp(q) for _, q in jobs]
wts = cnn.seeded_weights()
sp = cnn.SplitPredictor(wts)
logits, _ = sp.predict(host_in, 1, 8, W, H, len(jobs))
wts["fc_b"] = np.array([0.0, -float(np.median(logits[..., 1] - logits[..., 0]))], np.float32)
sp.close()
sp = cnn.SplitPredictor(wts)
logits, split = sp.predict(host_in, 1, 8, W, H, len(jobs))
print("predictor: %d CUs, kernel %.3f ms (rerun avg %.3f ms), split fraction %.3f" % (split.size, sp.last_kernel_ms()[0], sp.rerun_resident(10), split.mean()), flush=True)
sp.close()
hints = cnn.hints_from_split(W, H, split)
out = np.zeros(enc.num_ctus, fh.CTU_RESULT_DTYPE)
poc = 0
for chains in [int(a) for a in sys.argv[1:]] or [1, 2]:
    os.environ["FHC_HINTED_CHAINS"] = str(chains)
    t0 = time.perf_counter()
    enc.submit(host_in, sps, first_poc=poc, hints=hints)
    bits = 0
    for j in range(len(jobs)):
        enc.fetch_picture(poc + j, out)
        bits += int(out["totalBits"].sum(dtype=np.int64))
    ms = enc.last_kernel_ms()
    print("chains per block %d: kernel %.1f ms = %.2f fps (wall %.1f s), total bits %d" % (chains, ms, len(jobs) / ms * 1e3, time.perf_counter() - t0, bits), flush=True)
    poc += len(jobs)
enc.close()
