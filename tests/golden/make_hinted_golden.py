#!/usr/bin/env python3
"""Writes tests/golden/hinted_200x136_qp32.npz by RUNNING THE REAL REFERENCE carrying the CU-split hint hook
(oracle/_ref/TAppEncoderDumpHinted = HM 16.14 + the two-condition patch of INTEGRATION.md, built by
oracle/build_ref.py) with seeded random hints: every field of every finished CTU + the bitstream md5.  The file name
starts with "hinted_" so that the un-hinted golden tests skip it."""
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

w, h, n, qp, bd, sigma = 200, 136, 2, 32, 8, 12.0
d = tempfile.mkdtemp()
Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma)
write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
hints = np.random.default_rng(77).integers(0, 3, (n, T.n_ctus(w, h), 4)).astype(np.uint8)
hints.tofile(os.path.join(d, "hints.bin"))
bs, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, bit_depth=bd, exe="TAppEncoderDumpHinted",
                              env_extra={"FHC_HINTS": os.path.join(d, "hints.bin")})
pics = T.parse_dump(dump)
out = {"meta": np.array([w, h, n, qp, bd], np.int32), "sigma": np.array([sigma]), "hints": hints,
       "bitstream_md5": np.frombuffer(hashlib.md5(bs).digest(), np.uint8), "bitstream_len": np.array([len(bs)])}
for k, (s, ctus) in enumerate(pics):
    for f in T.FIELDS:
        a = ctus[f]
        out["p%d_%s" % (k, f)] = a.astype(np.int16) if f.startswith("coef") else a
np.savez_compressed(os.path.join(HERE, "hinted_200x136_qp32.npz"), **out)
print("bytes", len(bs))
