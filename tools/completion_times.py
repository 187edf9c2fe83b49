"""When do the pictures of the bench workload finish?  Submits the 256 jobs (64 frames x QP 22/27/32/37, 1920x1080) once and
polls the per-picture completion flags; prints first / median / last completion per QP class.  Usage (GPU box):
python tools/completion_times.py [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import fasthevc_b200 as fh  # noqa: E402
from fasthevc_b200.synth import synth_frames  # noqa: E402

QPS = (22, 27, 32, 37)


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    W, H = 1920, 1080
    jobs = [(f, q) for q in QPS for f in range(n_frames)]
    Y, U, V = synth_frames(W, H, n_frames)
    enc = fh.Encoder(W, H, 8, max_frames=len(jobs))
    host_in = enc.pinned_array((len(jobs), enc.frame_samples), np.uint8)
    packed = enc.pack_frames(Y, U, V)
    for j, (f, _) in enumerate(jobs):
        host_in[j] = packed[f]
    sps = [fh.slice_params_from_qp(q) for _, q in jobs]
    t0 = time.perf_counter()
    enc.submit(host_in, sps)
    done = np.zeros(len(jobs))
    left = set(range(len(jobs)))
    while left:
        for j in list(left):
            if enc.poll_picture(j):
                done[j] = time.perf_counter() - t0
                left.discard(j)
        time.sleep(0.02)
    print("kernel %.1f ms" % enc.last_kernel_ms())
    for k, q in enumerate(QPS):
        d = np.sort(done[k * n_frames:(k + 1) * n_frames])
        print("QP %d: first %.2f s, median %.2f s, last %.2f s" % (q, d[0], d[len(d) // 2], d[-1]))


if __name__ == "__main__":
    main()
