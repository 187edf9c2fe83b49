"""Dataset dump (fasthevc_b200/harp_dump.py) against the reference's serialiser and the oracle's
quad-trees.  The reference dumper needs OpenCV 3.4 C++ and cannot be built here: the JSON bytes are pinned by
the reference's own jsoncpp (tests/golden/harp_cu_json.json, made by tests/golden/make_harp_golden.py), the PNG bytes are
unpinned (decoded pixels are compared)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import fhc_testlib as T
from fasthevc_b200 import harp_dump as hd
from fasthevc_b200.synth import synth_frames

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "harp_cu_json.json")


def test_cu_json_bytes_match_reference_jsoncpp_golden():
    for g in json.load(open(GOLDEN)):
        rec = hd.cu_record(*g["case"])
        assert rec == g["fields"]
        assert hd.cu_json(rec) == g["reference_bytes"].encode()


@pytest.mark.skipif(not os.path.exists(os.path.join(T.ROOT, "oracle", "_ref", "harp_json_ref")), reason="oracle/_ref/harp_json_ref not built")
def test_cu_json_bytes_match_live_reference_jsoncpp():
    exe = os.path.join(T.ROOT, "oracle", "_ref", "harp_json_ref")
    rng = np.random.default_rng(7)
    for _ in range(20):
        d = int(rng.integers(0, 4))
        z = int(rng.integers(0, 1 << (2 * d))) * (256 >> (2 * d))       # a CU of depth d starts at a multiple of its partition count
        r = hd.cu_record(int(rng.integers(0, 64)), int(rng.integers(0, 510)), 30, z, d, int(rng.integers(0, 52)), "ds/F/CTU/CU%d" % z)
        args = [r["CTUIdx"], r["CTUPos"]["x"], r["CTUPos"]["y"], r["Pos"]["x"], r["Pos"]["y"], r["Width"], r["Depth"], r["AbsPartIdx"],
                r["RastPartIdx"], r["RelPos"]["x"], r["RelPos"]["y"], r["Mode"], r["QP"], r["POC"], r["imgSrc"][:-4]]
        ref = subprocess.run([exe] + [str(a) for a in args], capture_output=True, check=True).stdout
        assert hd.cu_json(r) == ref


def check_dump(out, poc, luma, ctus, qp, n_written):
    """Every leaf CU of the decided quad-trees has its record and its pixels; the records tile the picture exactly once."""
    import cv2
    h, w = luma.shape
    wc = (w + 63) // 64
    lum8 = np.clip(luma, 0, 255).astype(np.uint8)
    frame = cv2.imread(os.path.join(out, "F%d" % poc, "F%d-X-%d-Y-%d.png" % (poc, w, h)))
    assert frame.shape == (h, w, 3) and all(np.array_equal(frame[:, :, c], lum8) for c in range(3))
    cover = np.zeros((h, w), np.int32)
    n = 0
    for idx in range(len(ctus)):
        cx, cy = (idx % wc) * 64, (idx // wc) * 64
        cdir = os.path.join(out, "F%d" % poc, "CTU%d" % idx)
        ctu_img = cv2.imread(os.path.join(cdir, "CTU-%d-X-%d-Y-%d.png" % (idx, cx, cy)))
        assert np.array_equal(ctu_img[:, :, 0], lum8[cy:cy + 64, cx:cx + 64])
        depth = np.asarray(ctus[idx]["depth"])
        for name in sorted(os.listdir(cdir)):
            if not name.endswith(".json"):
                continue
            rec = json.load(open(os.path.join(cdir, name)))
            n += 1
            x, y, size, z = rec["Pos"]["x"], rec["Pos"]["y"], rec["Width"], rec["AbsPartIdx"]
            assert rec["Depth"] == int(depth[z]) and size == 64 >> rec["Depth"] and rec["Mode"] == "Intra" and rec["QP"] == qp and rec["POC"] == poc
            assert rec["CTUIdx"] == idx and (rec["CTUPos"]["x"], rec["CTUPos"]["y"]) == (cx, cy)
            assert name == "CU%d_%d.json" % (rec["RelPos"]["x"], rec["RelPos"]["y"]) and rec["imgSrc"] == os.path.join(cdir, name[:-5]) + ".png"
            assert rec["RastPartIdx"] == (rec["RelPos"]["y"] // 4) * 16 + rec["RelPos"]["x"] // 4 == int(hd.Z2R[z])
            img = cv2.imread(rec["imgSrc"])
            assert img.shape == (size, size, 3) and np.array_equal(img[:, :, 1], lum8[y:y + size, x:x + size])
            cover[y:y + size, x:x + size] += 1
    assert n == n_written and np.all(cover == 1)


@pytest.mark.parametrize("w,h,qp,bd", [(200, 136, 32, 8), (136, 72, 27, 10)])
def test_dump_from_oracle_results(w, h, qp, bd):
    Y, U, V = synth_frames(w, h, 1, bit_depth=bd, sigma=10.0, seed=77)
    ctus = T.oracle_compress(T.params_from_qp(w, h, bd, qp), Y[0], U[0], V[0])
    with tempfile.TemporaryDirectory() as out:
        n = hd.dump_picture(out, 5, Y[0], ctus, qp)
        check_dump(out, 5, Y[0], ctus, qp, n)


@pytest.mark.gpu
def test_dump_from_cuda_results_equals_dump_from_oracle():
    """The dump is a by-product of the fetched CTU results: produced from the CUDA path's records it is file for file the
    dump produced from the oracle's."""
    import fasthevc_b200 as fh
    w, h, qp = 264, 200, 32
    Y, U, V = synth_frames(w, h, 2, sigma=12.0, seed=99)
    enc = fh.Encoder(w, h, 8, max_frames=2)
    enc.submit(enc.pack_frames(Y, U, V), [fh.slice_params_from_qp(qp)] * 2)
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        for poc in range(2):
            got = enc.fetch_picture(poc)
            ref = T.oracle_compress(T.params_from_qp(w, h, 8, qp), Y[poc], U[poc], V[poc])
            n = hd.dump_picture(a, poc, Y[poc], got, qp)
            assert n == hd.dump_picture(b, poc, Y[poc], ref, qp)
            check_dump(a, poc, Y[poc], got, qp, n)
        for root, _, files in os.walk(a):
            for f in files:
                pa, pb = os.path.join(root, f), os.path.join(b + root[len(a):], f)
                ta, tb = open(pa, "rb").read(), open(pb, "rb").read()
                assert ta.replace(a.encode(), b"") == tb.replace(b.encode(), b""), pa
    enc.close()
