#!/usr/bin/env python3
"""Generates tests/golden/harp_cu_json.json: the bytes the reference's dataset dumper writes for a CU record
(Src_HARP/CShow_PredResiReco.h:79-99), produced by oracle/_ref/harp_json_ref = oracle/harp_json_ref.cpp linked with the
reference's own vendored jsoncpp (built by oracle/build_ref.py).  Run in the build container; the fixture is committed."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fasthevc_b200 import harp_dump as hd  # noqa: E402

CASES = [  # poc, ctuIdx, widthInCtus, absPartIdx, depth, qp, imgSrc (without .png)
    (0, 0, 7, 0, 0, 32, "dataset/CNNDataset_clip/F0/CTU0/CU0_0"),
    (3, 7, 7, 20, 2, 32, "dataset/CNNDataset_clip/F3/CTU7/CU24_0"),
    (63, 509, 30, 255, 3, 37, "/data/out dir/CNNDataset_x/F63/CTU509/CU60_60"),
    (5, 12, 30, 192, 1, 22, 'q"uo\\te/F5/CTU12/CU32_32'),
]


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "harp_json_ref")
    out = []
    for c in CASES:
        r = hd.cu_record(*c)
        args = [r["CTUIdx"], r["CTUPos"]["x"], r["CTUPos"]["y"], r["Pos"]["x"], r["Pos"]["y"], r["Width"], r["Depth"], r["AbsPartIdx"],
                r["RastPartIdx"], r["RelPos"]["x"], r["RelPos"]["y"], r["Mode"], r["QP"], r["POC"], r["imgSrc"][:-4]]
        text = subprocess.run([exe] + [str(a) for a in args], capture_output=True, check=True).stdout.decode()
        out.append({"case": list(c), "fields": r, "reference_bytes": text})
    json.dump(out, open(os.path.join(HERE, "harp_cu_json.json"), "w"), indent=1)
    print("wrote", len(out), "records")


if __name__ == "__main__":
    main()
