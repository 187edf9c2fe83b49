#!/usr/bin/env python3
"""Writes tests/golden/fullsize_<tag>.npz: one CRC-32 per CTU of the ORACLE's records (oracle/hm_intra.c) for the full-size
test pictures of fhc_testlib.FULLSIZE_CASES (1920x1080 at four QPs, 3840x2160 Main10; with and without wavefronts), so
that the GPU box does not spend minutes of host time on them.  The oracle itself is pinned against the real reference by
the small fixtures next to this file and the live runs of tests/test_cpu_oracle.py / test_wpp.py."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fhc_testlib as T  # noqa: E402


def job(arg):
    tag, k = arg
    return tag, k, T.ctu_crcs(T._oracle_job(T.fullsize_jobs(tag)[k]))


if __name__ == "__main__":
    T.build_oracle()
    args = [(tag, k) for tag, c in T.FULLSIZE_CASES.items() for k in range(len(c[5]))]
    args.sort(key=lambda a: -T.FULLSIZE_CASES[a[0]][0])            # the 2160p pictures first
    with mp.get_context("fork").Pool(min(len(args), os.cpu_count() or 4)) as pool:
        res = pool.map(job, args, chunksize=1)
    for tag, c in T.FULLSIZE_CASES.items():
        crc = np.stack([r[2] for r in sorted((r for r in res if r[0] == tag), key=lambda r: r[1])])
        np.savez_compressed(os.path.join(HERE, "fullsize_%s.npz" % tag), crc=crc)
        print(tag, crc.shape)
