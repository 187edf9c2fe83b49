"""Per-role cycle profile of the engine (FHC_PROFILE build, libfasthevc_cuda_prof.so): where a picture chain spends its time.
Usage (gpurun): FHC_LIBRARY=fasthevc_b200/libfasthevc_cuda_prof.so python tools/cycle_profile.py [W H N QP...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fasthevc_b200 as fh
from fasthevc_b200.synth import synth_frames

ROLES = ["3A 8x8 2Nx2N (+master)", "3B 8x8 NxN", "2  16x16", "1  32x32", "0  64x64"]
STAGES = ["rmd", "luma1", "luma2", "chroma", "syntax", "job", "wait", "jobs"]


def main():
    a = sys.argv[1:]
    w, h, n = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (416, 240, 1)
    qps = [int(q) for q in a[3:]] or [32]
    Y, U, V = synth_frames(w, h, min(n, 8), sigma=6.0)
    idx = np.arange(n) % Y.shape[0]
    enc = fh.Encoder(w, h, 8, max_frames=n, device=0)
    sps = [fh.slice_params_from_qp(qps[k % len(qps)]) for k in range(n)]
    enc.submit(enc.pack_frames(Y[idx], U[idx], V[idx]), sps)
    enc.fetch_picture(n - 1)
    ms = enc.last_kernel_ms()
    enc.lib.fhc_debug_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    ctus = enc.num_ctus
    print("%dx%d n=%d qps=%s: kernel %.1f ms, %.2f ms per CTU per chain" % (w, h, n, qps, ms, ms / ctus / max(1, -(-n // fh.pictures_in_flight(0)))))
    for pic in sorted(set([0, n // 2, n - 1])):
        out = np.zeros(64, np.int64)
        enc.lib.fhc_debug_profile(enc.ctx, pic, out.ctypes.data)
        tot = float(out[40]) or 1.0
        print("picture %d (qp %d): total %.1f Mcycles = %.2f Mcycles per CTU" % (pic, qps[pic % len(qps)], tot / 1e6, tot / 1e6 / ctus))
        print("  %-24s" % "role" + "".join("%9s" % s for s in STAGES))
        for r in range(5):
            row = out[r * 8:(r + 1) * 8]
            print("  %-24s" % ROLES[r] + "".join("%8.1f%%" % (100.0 * row[k] / tot) for k in range(7)) + "%9d" % row[7])
    enc.close()


if __name__ == "__main__":
    main()
