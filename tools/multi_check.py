"""Encodes N different small pictures in ONE submit and compares every picture with the oracle (debug helper).
usage: python tools/multi_check.py [n] [w] [h] [qp]"""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import fasthevc_b200 as fh
import fhc_testlib as T
from fasthevc_b200.synth import synth_frames
a = [int(x) for x in sys.argv[1:]]
n, w, h, qp = (a + [2, 64, 64, 32][len(a):])[:4]
Y, U, V = synth_frames(w, h, n, sigma=8.0)
if os.environ.get('FHC_SAME'):
    Y[:] = Y[0]; U[:] = U[0]; V[:] = V[0]
enc = fh.Encoder(w, h, 8, max_frames=n)
sp = [fh.slice_params_from_qp(qp)] * n
try:
    enc.submit(enc.pack_frames(Y, U, V), sp)
    enc.fetch_picture(0)
except Exception as ex:
    print('FAILED', str(ex)[-120:]); sys.exit(0)
p = T.params_from_qp(w, h, 8, qp)
nbad = 0
for k in range(n):
    got = enc.fetch_picture(k)
    ref = T.oracle_compress(p, Y[k], U[k], V[k])
    bad = T.compare_ctus(got, ref, "pic%d" % k)
    nbad += len(bad) > 0
    if bad:
        print(bad[:6])
print('pictures', n, 'mismatching', nbad)
enc.close()
