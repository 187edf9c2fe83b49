"""Worker of test_two_rank_sharding_gloo: frame groups over 2 gloo ranks, splice on rank 0."""
import hashlib
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fhc_testlib as T  # noqa: E402
from fasthevc_b200.sharding import frame_group, splice_in_poc_order  # noqa: E402
from fasthevc_b200.synth import synth_frames  # noqa: E402


def digest(rec):
    return hashlib.md5(rec.tobytes()).hexdigest()


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    w, h, n, qp = 72, 64, 3, 34
    first, last = frame_group(n, rank, ws)
    Y, U, V = synth_frames(w, h, last - first, sigma=10.0, first_frame=first)       # each rank generates only its group
    p = T.params_from_qp(w, h, 8, qp)
    mine = [digest(T.oracle_compress(p, Y[k], U[k], V[k])) for k in range(last - first)]
    gathered = [None] * ws
    dist.gather_object((first, mine), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        spliced = splice_in_poc_order(gathered)
        Yf, Uf, Vf = synth_frames(w, h, n, sigma=10.0)
        single = [digest(T.oracle_compress(p, Yf[k], Uf[k], Vf[k])) for k in range(n)]
        assert spliced == single, (spliced, single)
        print("SHARD_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
