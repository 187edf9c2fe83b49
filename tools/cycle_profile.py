import sys, time, ctypes as C; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import fasthevc_b200 as fh, fasthevc_b200.api as api
api.library_path = lambda: 'fasthevc_b200/libfasthevc_cuda_prof.so'
from fasthevc_b200.synth import synth_frames
names=['TOTAL','PAR','RDOQ','COEFBITS','REFS_MASTER','RMD_MASTER','TUBLOCK','BITSQT','CUCOPY','RMD_PAR','ENCODE','TASKS_LUMA','TASKS_CHROMA','LUMA','CHROMA']
def run(w,h,qp,sigma=6.0,n=1):
    Y,U,V=synth_frames(w,h,min(n,16),sigma=sigma)
    reps=(n+Y.shape[0]-1)//Y.shape[0]
    Y=np.concatenate([Y]*reps)[:n]; U=np.concatenate([U]*reps)[:n]; V=np.concatenate([V]*reps)[:n]
    enc=fh.Encoder(w,h,8,max_frames=n)
    sp=[fh.slice_params_from_qp(qp)]*n
    enc.submit(enc.pack_frames(Y,U,V),sp); o=enc.fetch_picture(0)
    cyc=(C.c_longlong*16)(); cnt=(C.c_longlong*16)()
    enc.lib.fhc_debug_profile.argtypes=[C.c_void_p,C.c_int,C.c_void_p,C.c_void_p]
    enc.lib.fhc_debug_profile(enc.ctx,0,cyc,cnt)
    tot=cyc[0]
    print('%dx%d qp %d sigma %.0f: total %.1f Mcycles, %d CTUs'%(w,h,qp,sigma,tot/1e6,enc.num_ctus))
    for i,nm in enumerate(names):
        if cnt[i]: print('  %-12s %6.1f%%  n=%8d  avg %8.0f cyc'%(nm,100.0*cyc[i]/tot,cnt[i],cyc[i]/cnt[i]))
    names2=['CHECK_D0','CHECK_D1','CHECK_D2','CHECK_D3_2N','CHECK_D3_NXN','LUMA2ND','CHROMA_BIG','LUMA1ST_BIG']
    enc.lib.fhc_debug_profile2.argtypes=[C.c_void_p,C.c_int,C.c_void_p,C.c_void_p]
    enc.lib.fhc_debug_profile2(enc.ctx,0,cyc,cnt)
    for i,nm in enumerate(names2):
        if cnt[i]: print('  %-12s %6.1f%%  n=%8d  avg %8.0f cyc'%(nm,100.0*cyc[i]/tot,cnt[i],cyc[i]/cnt[i]))
    enc.close()
run(192,128,32,n=int(sys.argv[1]) if len(sys.argv)>1 else 1)
