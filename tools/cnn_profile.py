#!/usr/bin/env python3
"""One launch of the CNN split predictor over 64 1080p pictures (126,720 CUs) for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:fhc_cnn -c 1 -o gpurun_out/cnn python tools/cnn_profile.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fasthevc_b200 as fh
from fasthevc_b200 import cnn
from fasthevc_b200.synth import synth_frames

W, H, N = 1920, 1080, 64
Y, U, V = synth_frames(W, H, 8)
idx = [k % 8 for k in range(N)]
yuv = fh.Encoder.pack_frames(Y[idx], U[idx], V[idx])
sp = cnn.SplitPredictor(cnn.seeded_weights())
logits, split = sp.predict(yuv, 1, 8, W, H, N)
ms, cus = sp.last_kernel_ms()
print("predictor: %d CUs, kernel %.3f ms, %.1f M CU/s, %.1f TFLOP/s" % (cus, ms, cus / ms / 1e3, 5.62e6 * cus / ms / 1e9))
sp.close()
