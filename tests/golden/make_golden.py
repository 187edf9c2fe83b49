#!/usr/bin/env python3
"""Generates the golden per-CTU fixtures in this directory by RUNNING THE REAL REFERENCE
(oracle/_ref/TAppEncoderDump, built by oracle/build_ref.py from /root/reference) on seeded
synthetic clips.  Only needed when fixtures are regenerated; the .npz files are committed.

Each fixture holds: the clip parameters, the slice-level lambdas the reference used, every
field of every finished CTU (metadata, coefficients, reconstruction, cost/dist/bits/bins),
and the md5 of the reference bitstream."""
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

CASES = [  # name, w, h, frames, qp, bit depth, sigma
    ("c1_416x240_qp32", 416, 240, 1, 32, 8, 6.0),
    ("tiny_136x72_qp32", 136, 72, 2, 32, 8, 12.0),
    ("busy_200x136_qp22", 200, 136, 1, 22, 8, 20.0),
    ("main10_136x72_qp27", 136, 72, 1, 27, 10, 12.0),
    ("smooth_192x128_qp37", 192, 128, 1, 37, 8, 3.0),
]


def main():
    for name, w, h, n, qp, bd, sigma in CASES:
        d = tempfile.mkdtemp()
        Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma)
        write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
        bs, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, n, qp, d, bit_depth=bd)
        pics = T.parse_dump(dump)
        assert len(pics) == n
        out = {"meta": np.array([w, h, n, qp, bd], np.int32), "sigma": np.array([sigma]),
               "bitstream_md5": np.frombuffer(hashlib.md5(bs).digest(), np.uint8), "bitstream_len": np.array([len(bs)])}
        for k, (s, ctus) in enumerate(pics):
            out["slice%d" % k] = np.array([s["lambda"], s["sqrtLambda"], s["distWeightCb"], s["distWeightCr"]] + list(s["rdoqLambda"]))
            for f in T.FIELDS:
                a = ctus[f]
                if f.startswith("coef"):
                    a = a.astype(np.int16)
                out["p%d_%s" % (k, f)] = a
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "bytes", len(bs), "->", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
