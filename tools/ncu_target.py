import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames
w,h,qp=64,64,32
Y,U,V=synth_frames(w,h,1,sigma=8.0)
enc=fh.Encoder(w,h,8,max_frames=1)
sp=[fh.slice_params_from_qp(qp)]
enc.submit(enc.pack_frames(Y,U,V),sp); o=enc.fetch_picture(0)
print('bits',int(o['totalBits'].sum()))
enc.close()
