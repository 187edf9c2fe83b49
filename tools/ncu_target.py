"""ncu target: one submit of N small pictures (default 1184 = 8 warps x 148 SMs of 64x64, one CTU each) so that the
capture shows the kernel under full-chip load.  usage: ncu ... python tools/ncu_target.py [n] [w] [h] [qp]"""
import os
import sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames
a = [int(x) for x in sys.argv[1:]]
n, w, h, qp = (a + [1184, 64, 64, 32][len(a):])[:4]
Y, U, V = synth_frames(w, h, min(n, 8), sigma=8.0)
reps = (n + Y.shape[0] - 1) // Y.shape[0]
Y = np.concatenate([Y] * reps)[:n]; U = np.concatenate([U] * reps)[:n]; V = np.concatenate([V] * reps)[:n]
enc = fh.Encoder(w, h, 8, max_frames=n)
sp = [fh.slice_params_from_qp(qp)] * n
enc.submit(enc.pack_frames(Y, U, V), sp)
o = enc.fetch_picture(0)
print('bits', int(o['totalBits'].sum()))
enc.close()
