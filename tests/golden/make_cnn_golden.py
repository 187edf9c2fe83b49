"""Writes tests/golden/cnn_split_416x240.npz: the logits OpenCV's fp32 dnn module produces for the seeded stand-in
network (oracle/cnn_split.py: Darknet export -> cv2.dnn.readNetFromDarknet) on the depth-1 CUs of two synthetic
416x240 pictures.  PARITY UNPINNED in the reference itself (no inference code or weights exist there):
this fixture pins the numpy restatement and the CUDA kernel to an independent fp32 implementation (OpenCV)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cnn_split as cs  # noqa: E402
from fasthevc_b200.synth import synth_frames  # noqa: E402

SEED, FRAMES = 20240, 2
w = cs.make_weights(SEED)
Y, U, V = synth_frames(416, 240, FRAMES, sigma=12.0)
patches = np.concatenate([cs.gather_patches(Y[i]) for i in range(FRAMES)])
with tempfile.TemporaryDirectory() as d:
    lg = cs.forward_opencv(w, patches, d)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cnn_split_416x240.npz"), seed=SEED, frames=FRAMES, patches=patches,
                    logits_opencv=lg, split=cs.split_decision(lg))
print("wrote %d CUs, %d split" % (len(patches), int(cs.split_decision(lg).sum())))
