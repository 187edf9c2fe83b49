"""Exact work counts of the hot path per CTU: samples through the TU pipeline by TU size, samples of
the rough mode decision and CABAC bins counted, from the host emulation of the kernel source (bit-exact with the
reference), per QP on one synthetic 1920x1080 picture of bench.py's clip.  TEST/TOOLS ONLY (uses tests/_build).
usage: python tools/work_count.py [width height]  ->  markdown table on stdout"""
import ctypes as C
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import fhc_testlib as T
from fasthevc_b200.synth import synth_frames

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
emu = C.CDLL(T.build_emu())
emu.fhc_emu_counters.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
WEIGHT = [94, 104, 125, 168]          # integer-equivalent ops per sample of the TU pipeline by TU size
Y, U, V = synth_frames(W, H, 1)
nctu = T.n_ctus(W, H)
print("| QP | S_rmd / CTU | S_tq luma 4/8/16/32 / CTU | S_tq chroma 4/8/16 / CTU | TU codings / CTU | bins / CTU | W = 13 S_rmd + sum w S_tq + 30 bins |")
print("|---|---|---|---|---|---|---|")
for qp in (22, 27, 32, 37):
    out = (C.c_ulonglong * 18)()
    emu.fhc_emu_counters(out, 1)
    T.emu_compress(T.params_from_qp(W, H, 8, qp), Y[0], U[0], V[0])
    emu.fhc_emu_counters(out, 1)
    v = [float(x) / nctu for x in out]
    w = 13 * v[16] + sum(WEIGHT[k] * (v[k] + v[4 + k]) for k in range(4)) + 30 * v[17]
    print("| %d | %.0f | %.0f / %.0f / %.0f / %.0f | %.0f / %.0f / %.0f | %.0f | %.0f | %.1f M |" % (
        qp, v[16], v[0], v[1], v[2], v[3], v[4], v[5], v[6], sum(v[8:16]), v[17], w / 1e6))
