import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import gpu_check as G
import fasthevc_b200 as fh
print("in flight", fh.pictures_in_flight(0))
for n in (74, 148, 296):
    G.perf(416, 240, n, [22, 27, 32, 37])
