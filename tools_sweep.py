import sys, os, subprocess
cfgs={'A':(1,296),'C':(1,148),'D':(1,296),'E':(1,444)}
for k,(a,b) in cfgs.items():
    env=dict(os.environ); env['FHC_LIBRARY']='fasthevc_b200/libfhc_%s.so'%k
    r=subprocess.run([sys.executable,'tools_time2.py',str(a),str(b)],env=env,capture_output=True,text=True,timeout=400)
    print('== cfg',k); print(r.stdout[-400:]); print(r.stderr[-300:])
