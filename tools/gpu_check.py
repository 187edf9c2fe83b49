"""Quick GPU check of the CUDA path (gpurun): parity on a few small cases against the oracle, then throughput probes.
Usage: python tools/gpu_check.py [parity] [perf] [perf1080]"""
import sys
import time
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fasthevc_b200 as fh
import fhc_testlib as T
from fasthevc_b200.synth import synth_frames


def run(w, h, bd, qps, Y, U, V):
    n = Y.shape[0]
    enc = fh.Encoder(w, h, bd, max_frames=n, device=0)
    sps = [T.slice_params_from_hm(T.params_from_qp(w, h, bd, qps[k % len(qps)])) for k in range(n)]
    t0 = time.time()
    enc.submit(enc.pack_frames(Y, U, V), sps)
    out = [enc.fetch_picture(k) for k in range(n)]
    dt = time.time() - t0
    ms = enc.last_kernel_ms()
    enc.close()
    return out, dt, ms


def parity(tiny=False):
    cases = [(8, 8, 1, 32, 8, 8.0), (64, 64, 1, 30, 8, 8.0)] if tiny else [(8, 8, 1, 32, 8, 8.0), (64, 64, 1, 30, 8, 8.0), (136, 72, 3, 32, 8, 10.0), (416, 240, 1, 22, 8, 20.0),
             (416, 240, 2, 32, 8, 6.0), (200, 136, 2, 27, 10, 12.0), (264, 200, 1, 37, 8, 12.0)]
    ok = True
    for (w, h, n, qp, bd, sigma) in cases:
        Y, U, V = synth_frames(w, h, n, bit_depth=bd, sigma=sigma, seed=4321)
        got, dt, ms = run(w, h, bd, [qp], Y, U, V)
        p = T.params_from_qp(w, h, bd, qp)
        bad = []
        for k in range(n):
            bad += T.compare_ctus(got[k], T.oracle_compress(p, Y[k], U[k], V[k]), "pic%d" % k)
        print("parity %dx%d n=%d qp=%d bd=%d: %s (kernel %.1f ms)" % (w, h, n, qp, bd, "OK" if not bad else "MISMATCH %d %s" % (len(bad), bad[:6]), ms), flush=True)
        ok = ok and not bad
    return ok


def perf(w, h, n, qps, distinct=8):
    Yd, Ud, Vd = synth_frames(w, h, distinct, sigma=6.0)
    idx = np.arange(n) % distinct
    got, dt, ms = run(w, h, 8, qps, Yd[idx], Ud[idx], Vd[idx])
    ctus = T.n_ctus(w, h) * n
    print("perf %dx%d n=%d qps=%s: kernel %.1f ms, wall %.2f s -> %.0f CTU/s, %.2f fps (kernel), %.2f ms per CTU per block chain"
          % (w, h, n, qps, ms, dt, ctus / (ms / 1e3), n / (ms / 1e3), ms / T.n_ctus(w, h) / max(1, -(-n // fh.pictures_in_flight(0)))), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "perf"]
    print("pictures in flight:", fh.pictures_in_flight(0), flush=True)
    if "tiny" in what:
        parity(True)
    if "parity" in what:
        parity()
    if "perf" in what:
        perf(416, 240, 1, [32])
        perf(416, 240, 32, [32])
        perf(416, 240, 296, [32])
        perf(416, 240, 296, [22, 27, 32, 37])
    if "ncucase" in what:
        perf(136, 72, 296, [22, 27, 32, 37])
    if "ncucase2" in what:
        perf(416, 240, 296, [22, 27, 32, 37])
    if "perf256" in what:
        perf(416, 240, 256, [22, 27, 32, 37])
        # QP-major order as in bench.py
        import fasthevc_b200 as fh2
    if "ab" in what:                      # three repetitions of both loaded cases (library comparisons on the same box via FHC_LIBRARY)
        for _ in range(3):
            perf(416, 240, 296, [22, 27, 32, 37])
            perf(416, 240, 256, [22, 27, 32, 37])
    if "lonecase" in what:                # every picture alone in its block (ncu target for the lone-chain profile)
        perf(416, 240, 148, [27])
    if "lone" in what:                    # one / two / three pictures per SM
        for n in (148, 296, 444):
            perf(416, 240, n, [22, 27, 32, 37])
    if "perf2160" in what:                # BASELINE configs[3]: 3840x2160 Main10, 32 frames at QP 27 (every picture alone in its block)
        Yd, Ud, Vd = synth_frames(3840, 2160, 4, bit_depth=10, sigma=6.0)
        idx = np.arange(32) % 4
        got, dt, ms = run(3840, 2160, 10, [27], Yd[idx], Ud[idx], Vd[idx])
        print("perf 3840x2160 Main10 n=32 qp=27: kernel %.1f ms -> %.2f fps, %.0f CTU/s" % (ms, 32 / (ms / 1e3), 32 * T.n_ctus(3840, 2160) / (ms / 1e3)), flush=True)
    if "perf1080" in what:
        perf(1920, 1080, 8, [22, 27, 32, 37])
        perf(1920, 1080, 256, [22, 27, 32, 37], distinct=16)
