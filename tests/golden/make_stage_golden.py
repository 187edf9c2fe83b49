#!/usr/bin/env python3
"""Generates tests/golden/stages_64x64_qp32.npz: per-stage vectors of ONE CTU from the real reference, printed by HM's own
debug switches (TypeDef.h:58-59: DEBUG_INTRA_SEARCH_COSTS, DEBUG_TRANSFORM_AND_QUANTISE) in oracle/_ref/TAppEncoderStages
(built by oracle/build_ref.py).  In evaluation order:

  R  first-pass candidate of a PU: mode, SATD, mode bits               (TEncSearch.cpp:2290)  11,935 records = 341 PUs x 35 modes
  P  full-RD cost of a candidate (as printed: 6 significant digits)     (TEncSearch.cpp:2370)
  T  a TU on its way through the codec: size, component, CRC-32 of the residual, of the transform output, of the
     quantiser (RDOQ) output and - when the TU is coded - of the dequantiser and inverse-transform outputs
     (TComTrQuant.cpp:1496-1526, :1603-1660); the first records are also kept in full

tests/test_cpu_oracle.py::test_oracle_stage_trace_matches_reference replays the oracle with its trace on and compares.
Run in the build container; the fixture is committed."""
import os
import re
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fasthevc_b200.synth import synth_frames, write_yuv  # noqa: E402

W = H = 64
QP, SIGMA, SEED = 32, 8.0, 4321
KEEP_FULL = 6


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a, dtype="<i4").tobytes())


def parse(text):
    tags, rmd, p2, tu, full = [], [], [], [], []
    lines = text.split("\n")
    i = 0
    head = re.compile(r"^(\d+): (\d+)x(\d+) channel (\d) TU (.*)$")

    def block(i, n):
        rows = [np.array(lines[i + 1 + r].split(), dtype=np.int64) for r in range(n)]
        return np.stack(rows).astype(np.int32), i + 1 + n

    while i < len(lines):
        l = lines[i]
        if l.startswith("1st pass mode"):
            m = re.match(r"1st pass mode (\d+) SAD = (\d+), mode bits = (\d+), cost = (.*)", l)
            tags.append(ord("R")); rmd.append([int(m.group(1)), int(m.group(2)), int(m.group(3))])
            i += 1
        elif l.startswith("2nd pass"):
            m = re.match(r"2nd pass \[luma,chroma\] mode \[(\d+),(\d+)\] cost = (.*)", l)
            tags.append(ord("P")); p2.append([float(m.group(1)), float(m.group(3))])
            i += 1
        elif head.match(l):
            m = head.match(l)
            n, comp, what = int(m.group(2)), int(m.group(4)), m.group(5)
            b, i = block(i, n)
            if what == "at input to transform":
                tags.append(ord("T")); tu.append([n, comp, crc(b), 0, 0, 0, 0])
                if len(full) < KEEP_FULL:
                    full.append({"n": n, "comp": comp, "resi": b})
            elif what == "between transform and quantiser":
                tu[-1][3] = crc(b)
                if len(full) <= KEEP_FULL and "tr" not in full[-1] and len(tu) <= KEEP_FULL:
                    full[-1]["tr"] = b
            elif what == "at output of quantiser":
                tu[-1][4] = crc(b)
                if "q" not in full[-1] and len(tu) <= KEEP_FULL:
                    full[-1]["q"] = b
            elif what == "between dequantiser and inverse-transform":
                tu[-1][5] = crc(b)
            elif what == "at output of inverse-transform":
                tu[-1][6] = crc(b)
        else:
            i += 1
    return np.array(tags, np.uint8), np.array(rmd, np.int32), np.array(p2, np.float64), np.array(tu, np.int64), full


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderStages")
    cfg = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_intra_main.cfg")
    Y, U, V = synth_frames(W, H, 1, sigma=SIGMA, seed=SEED)
    with tempfile.TemporaryDirectory() as d:
        write_yuv(os.path.join(d, "clip.yuv"), Y, U, V)
        r = subprocess.run([exe, "-c", cfg, "-i", "clip.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "1", "-q", str(QP), "-b", "o.bin", "-o", ""],
                           capture_output=True, text=True, cwd=d, check=True)
    tags, rmd, p2, tu, full = parse(r.stdout)
    out = {"params": np.array([W, H, QP, 8, SEED], np.int32), "sigma": np.array([SIGMA]), "tags": tags, "rmd": rmd, "pass2": p2, "tu": tu}
    for k, f in enumerate(full):
        out["full%d_hdr" % k] = np.array([f["n"], f["comp"]], np.int32)
        for key in ("resi", "tr", "q"):
            out["full%d_%s" % (k, key)] = f[key]
    path = os.path.join(HERE, "stages_%dx%d_qp%d.npz" % (W, H, QP))
    np.savez_compressed(path, **out)
    print("wrote %s: %d R, %d P, %d T records, %d bytes" % (path, len(rmd), len(p2), len(tu), os.path.getsize(path)))


if __name__ == "__main__":
    main()
