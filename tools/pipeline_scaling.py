#!/usr/bin/env python3
"""Host side of the pipeline (fasthevc_b200/pipeline.py) on this machine's cores, device cost taken out (build container, CPU
only): the device's results for a 1080p clip are prepared once by the oracles (mode decision, deblocking, SAO statistics),
then the clip is encoded with 1, 2, 4, ... host encoder processes reading them from the spool.  What it shows: how many
pictures per second the unchanged host tail delivers per process count - the figure that has to exceed the device's.
usage: python tools/pipeline_scaling.py [pictures] [qp] [workers ...]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fhc_testlib as T  # noqa: E402
from fasthevc_b200 import api, pipeline  # noqa: E402
from fasthevc_b200.synth import synth_frames, write_yuv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
qp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
worker_counts = [int(a) for a in sys.argv[3:]] or [1, 2, 4, 8]
w, h, distinct = 1920, 1080, 4
d = tempfile.mkdtemp()
Y, U, V = synth_frames(w, h, distinct, sigma=6.0)
idx = [k % distinct for k in range(n)]
yuv = os.path.join(d, "s.yuv")
write_yuv(yuv, Y[idx], U[idx], V[idx])
t0 = time.perf_counter()
recs = T.oracle_compress_parallel([((w, h, 8, qp), Y[k], U[k], V[k]) for k in range(distinct)])
planes, stats = [], []
for k in range(distinct):
    dy, du, dv = T.oracle_deblock(w, h, 8, qp, recs[k])
    planes.append(np.concatenate([dy.reshape(-1), du.reshape(-1), dv.reshape(-1)]).astype(np.uint8))
    s = T.oracle_sao_stats(w, h, 8, (dy, du, dv), (Y[k], U[k], V[k]))
    c = np.zeros((T.n_ctus(w, h), 3), api.SAO_STATS_DTYPE)
    c["eoDiff"], c["eoCount"] = s["diff"][:, :, :4, :5], s["count"][:, :, :4, :5]
    c["boDiff"], c["boCount"] = s["diff"][:, :, 4, :], s["count"][:, :, 4, :]
    stats.append(c)
print("oracles: %d distinct pictures in %.1f s" % (distinct, time.perf_counter() - t0))
spool = os.path.join(d, "spool")
os.makedirs(spool)
sp = api.slice_params_from_qp(qp)
for p in range(n):
    pipeline.write_spool_picture(spool, p, w, h, 8, False, sp, recs[p % distinct].astype(api.CTU_RESULT_DTYPE), planes[p % distinct], stats[p % distinct])


class Done:                                   # the device side has nothing left to do: everything is in the spool
    class Seq:
        width, height, bitDepth = w, h, 8
    seq = Seq()

    def submit(self, *a, **k):
        pass

    def deblock_pictures_stats(self, first, m):
        return [None] * m, [None] * m

    def fetch_picture(self, poc):
        return None

    def close(self):
        pass


pipeline.build_spool_library()
orig_write = pipeline.write_spool_picture
pipeline.write_spool_picture = lambda *a, **k: None      # keep the prepared files
exe, cfg = os.path.join(T.REF_DIR, "TAppEncoderFHC"), os.path.join(T.REF_DIR, "cfg", "encoder_intra_main.cfg")
ref = None
for workers in worker_counts:
    out = os.path.join(d, "w%d.bin" % workers)
    info = pipeline.encode(exe, cfg, yuv, w, h, n, qp, out, workers, batch=n, backend=Done(), spool_dir=spool)
    data = open(out, "rb").read()
    ref = ref or data
    print("%2d host processes: %d pictures 1080p QP %d in %.2f s = %.1f fps (bitstream %d bytes, identical to the 1-process one: %s)"
          % (info["workers"], n, qp, info["wall_s"], info["fps"], len(data), data == ref), flush=True)
ok, bad = T.decoder_md5_ok(out, d)
print("decoder MD5 ok %d/%d" % (ok, n))
