#!/usr/bin/env python3
"""Writes tests/golden/wpp_264x200_qp32.npz by RUNNING THE REAL REFERENCE (oracle/_ref/TAppEncoderDump) with
--WaveFrontSynchro=1 on a seeded synthetic clip (5 x 4 CTUs, boundary CTUs on both edges): every field of every finished
CTU + the bitstream md5.  The name starts with "wpp_" so that the plain golden tests skip it."""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fhc_testlib as T  # noqa: E402
from fasthevc_b200.synth import synth_frames, write_yuv  # noqa: E402

w, h, n, qp, bd, sigma = 264, 200, 1, 32, 8, 12.0
d = tempfile.mkdtemp()
Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma)
write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
bs, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, bit_depth=bd, extra=["--WaveFrontSynchro=1"])
pics = T.parse_dump(dump)
out = {"meta": np.array([w, h, n, qp, bd], np.int32), "sigma": np.array([sigma]),
       "bitstream_md5": np.frombuffer(hashlib.md5(bs).digest(), np.uint8), "bitstream_len": np.array([len(bs)])}
for k, (s, ctus) in enumerate(pics):
    for f in T.FIELDS:
        a = ctus[f]
        out["p%d_%s" % (k, f)] = a.astype(np.int16) if f.startswith("coef") else a
np.savez_compressed(os.path.join(HERE, "wpp_264x200_qp32.npz"), **out)
print("bytes", len(bs))
