#!/usr/bin/env python3
"""Writes tests/golden/deblock_*.npz by RUNNING THE REAL REFERENCE (oracle/_ref/TAppEncoderDump) with --SAO=0: the
reconstruction file then holds the pictures exactly as TComLoopFilter::loopFilterPic leaves them.  Each fixture holds the
deblocked planes and what the deblocking oracle / kernel need as input (CU depth, transform index, pre-deblocking
reconstruction of every CTU, taken from the reference's per-CTU dump), plus the reference's SAO statistics of the deblocked
picture (m_statData after TEncSampleAdaptiveOffset::getStatistics, from a second run with SAO on)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fhc_testlib as T  # noqa: E402
from fasthevc_b200.synth import synth_frames, write_yuv  # noqa: E402

for name, w, h, qp, bd, sigma in (("deblock_264x200_qp34", 264, 200, 34, 8, 5.0), ("deblock_main10_136x72_qp37", 136, 72, 37, 10, 6.0)):
    d = tempfile.mkdtemp()
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=sigma)
    write_yuv(os.path.join(d, "s.yuv"), Y, U, V)
    bs, dump, _ = T.run_reference(os.path.join(d, "s.yuv"), w, h, 1, qp, d, bit_depth=bd, extra=["--SAO=0"])
    s, ctus = T.parse_dump(dump)[0]
    rec = np.fromfile(os.path.join(d, "rec.yuv"), dtype=np.uint8 if bd == 8 else np.uint16)
    # the reference's SAO statistics of the same picture (SAO on: m_statData right after getStatistics, FHC_SAO_DUMP)
    T.run_reference(os.path.join(d, "s.yuv"), w, h, 1, qp, d, bit_depth=bd, env_extra={"FHC_SAO_DUMP": os.path.join(d, "sao.bin")}, tag="_sao")
    sao = np.fromfile(os.path.join(d, "sao.bin"), dtype=T.SAO_HM_DT).reshape(T.n_ctus(w, h), 3, 5)
    assert np.abs(sao["diff"]).max() < 2 ** 31 and sao["count"].max() < 2 ** 31
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array([w, h, 1, qp, bd], np.int32), sigma=np.array([sigma]),
                        depth=ctus["depth"], trIdx=ctus["trIdx"], recoY=ctus["recoY"], recoCb=ctus["recoCb"], recoCr=ctus["recoCr"],
                        deblocked=rec, sao_diff=sao["diff"].astype(np.int32), sao_count=sao["count"].astype(np.int32))
    print(name, len(rec))
