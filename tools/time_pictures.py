import sys, time; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames
def run(w,h,n,qp):
    Y,U,V=synth_frames(w,h,min(n,int(os.environ.get('FHC_DISTINCT','4'))))
    reps=(n+Y.shape[0]-1)//Y.shape[0]
    Y=np.concatenate([Y]*reps)[:n]; U=np.concatenate([U]*reps)[:n]; V=np.concatenate([V]*reps)[:n]
    enc=fh.Encoder(w,h,8,max_frames=n)
    qps=[int(q) for q in os.environ.get('FHC_QPS',str(qp)).split(',')]
    sp=[fh.slice_params_from_qp(qps[(i*len(qps))//n]) for i in range(n)]   # QP-major, like bench.py
    enc.submit(enc.pack_frames(Y,U,V),sp); o=enc.fetch_picture(0)
    ms=enc.last_kernel_ms()
    print('%dx%d n=%d qp=%d: kernel %.1f ms -> %.2f fps, %.1f CTU/s'%(w,h,n,qp,ms,n/(ms/1e3),n*enc.num_ctus/(ms/1e3)), flush=True)
    enc.close()
for n in [int(a) for a in sys.argv[1:]] or [1,148,296]:
    run(416,240,n,32)
