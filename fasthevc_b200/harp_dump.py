"""Training-set dump of the reference's dataset generator as a by-product of the fetched CTU results.

The reference hooks a HARP visualiser into the encoder (TEncGOP.cpp:1562 -> Src_HARP/CShow_PredResiReco.h) that, for every
picture, writes

    <out>/F<poc>/F<poc>-X-<W>-Y-<H>.png                       the original luma picture as a 3-channel 8-bit PNG (:166-172)
    <out>/F<poc>/CTU<i>/CTU-<i>-X-<x>-Y-<y>.png               the CTU's part of it, cropped at the picture edge (:190-201)
    <out>/F<poc>/CTU<i>/CU<relx>_<rely>.png / .json           every leaf CU of the decided quad-tree (:213-252, CuData :52-99)

The JSON record holds CTUIdx, CTUPos, Pos, Width, Depth, AbsPartIdx, RastPartIdx, RelPos, Mode, QP, POC, imgSrc
(get_CU_Internal / get_CU_FIN, Src_HARP/CShow.h:100-182) and is serialised by the reference's vendored jsoncpp
(`file << root << endl`: keys in alphabetical order, tab indentation, " : " separators) - `cu_json` reproduces those
bytes (pinned by tests/golden/harp_cu_json.json, generated with the reference's own jsoncpp).  The PNG bytes depend on
the OpenCV / libpng build and are NOT pinned (parity unpinned for the images; the tests compare decoded
pixels).  Here the inputs are the original luma plane and the `fhc_ctu_result` records of a picture, i.e. what
`Encoder.fetch_picture` returns: no second pass over the encoder state is needed.

The quad-tree walk follows xDrawCU (:213-252): a node whose top-left partition has a larger depth is split, children
whose top-left sample lies outside the picture are skipped.
"""
import json
import os

import numpy as np

CTU_DIM, SU_DIM = 64, 4


def _z2r():
    """g_auiZscanToRaster for a 64x64 CTU with 4x4 partitions (TComRom.cpp initZscanToRaster)."""
    out = np.zeros(256, np.int32)
    for z in range(256):
        x = y = 0
        for b in range(4):
            x |= ((z >> (2 * b)) & 1) << b
            y |= ((z >> (2 * b + 1)) & 1) << b
        out[z] = y * 16 + x
    return out


Z2R = _z2r()


def leaf_cus(depth, ctu_x, ctu_y, width, height):
    """Leaf CUs of one CTU in the order xDrawCU visits them: list of (absPartIdx, depth)."""
    out = []

    def walk(z, d):
        if int(depth[z]) > d:
            step = (256 >> (2 * d)) >> 2
            for k in range(4):
                zz = z + k * step
                r = int(Z2R[zz])
                if ctu_x + (r % 16) * SU_DIM < width and ctu_y + (r // 16) * SU_DIM < height:
                    walk(zz, d + 1)
            return
        out.append((z, int(depth[z])))

    walk(0, 0)
    return out


def cu_record(poc, ctu_idx, width_in_ctus, abs_part_idx, depth, qp, img_src, mode="Intra"):
    """The fields of sCU / CuData for one leaf CU (Src_HARP/CShow.h:139-182)."""
    rast = int(Z2R[abs_part_idx])
    ctu_pos = ((ctu_idx % width_in_ctus) * CTU_DIM, (ctu_idx // width_in_ctus) * CTU_DIM)
    rel = ((rast % (CTU_DIM // SU_DIM)) * SU_DIM, (rast // (CTU_DIM // SU_DIM)) * SU_DIM)
    return {"CTUIdx": int(ctu_idx), "CTUPos": {"x": ctu_pos[0], "y": ctu_pos[1]}, "Pos": {"x": ctu_pos[0] + rel[0], "y": ctu_pos[1] + rel[1]},
            "Width": CTU_DIM >> depth, "Depth": int(depth), "AbsPartIdx": int(abs_part_idx), "RastPartIdx": rast,
            "RelPos": {"x": rel[0], "y": rel[1]}, "Mode": mode, "QP": int(qp), "POC": int(poc), "imgSrc": img_src + ".png"}


def cu_json(rec):
    """Bytes of `file << root << endl` for a CuData record (jsoncpp StreamWriterBuilder defaults)."""
    def value(v, indent):
        if isinstance(v, dict):
            inner = indent + "\t"
            body = ",\n".join('%s%s : %s' % (inner, json.dumps(k), value(v[k], inner)) for k in sorted(v))
            return "\n%s{\n%s\n%s}" % (indent, body, indent)
        return json.dumps(v)

    body = ",\n".join('\t%s : %s' % (json.dumps(k), value(rec[k], "\t")) for k in sorted(rec))
    return ("{\n%s\n}\n" % body).encode()


def _to_8uc3(plane):
    """convertToCV_8UC3 (Src_HARP/CHelper.h:267-280): grey -> three equal channels, saturating conversion to 8 bits."""
    p = np.asarray(plane)
    if p.dtype != np.uint8:
        p = np.clip(p, 0, 255).astype(np.uint8)
    return np.repeat(p[:, :, None], 3, axis=2)


def dump_picture(out_dir, poc, luma, ctu_results, qp, write_png=None):
    """Writes the dump of one picture; returns the number of CU records written.

    luma: (H, W) original luma plane; ctu_results: the picture's records (structured array with `depth`, `predMode`, or
    a sequence of mappings); qp: slice QP (pcCU->getQP: no delta QP in the all-intra cfg)."""
    if write_png is None:
        import cv2

        def write_png(path, img):
            if not cv2.imwrite(path, img, [cv2.IMWRITE_PNG_COMPRESSION, 0]):
                raise IOError("cannot write " + path)
    luma = np.asarray(luma)
    h, w = luma.shape
    wc, hc = (w + 63) // 64, (h + 63) // 64
    if len(ctu_results) != wc * hc:
        raise ValueError("expected %d CTU records, got %d" % (wc * hc, len(ctu_results)))
    org = _to_8uc3(luma)
    fdir = os.path.join(out_dir, "F%d" % poc)
    os.makedirs(fdir, exist_ok=True)
    write_png(os.path.join(fdir, "F%d-X-%d-Y-%d.png" % (poc, w, h)), org)
    n = 0
    for idx in range(wc * hc):
        cx, cy = (idx % wc) * 64, (idx // wc) * 64
        cdir = os.path.join(fdir, "CTU%d" % idx)
        os.makedirs(cdir, exist_ok=True)
        write_png(os.path.join(cdir, "CTU-%d-X-%d-Y-%d.png" % (idx, cx, cy)), org[cy:min(cy + 64, h), cx:min(cx + 64, w)])
        depth = np.asarray(ctu_results[idx]["depth"])
        for z, d in leaf_cus(depth, cx, cy, w, h):
            r = int(Z2R[z])
            rx, ry = (r % 16) * 4, (r // 16) * 4
            name = os.path.join(cdir, "CU%d_%d" % (rx, ry))
            size = 64 >> d
            write_png(name + ".png", org[cy + ry:cy + ry + size, cx + rx:cx + rx + size])
            pm = int(np.asarray(ctu_results[idx]["predMode"])[z])
            if pm != 1:
                raise ValueError("CTU %d partition %d: prediction mode %d (only intra CUs exist in the all-intra path)" % (idx, z, pm))
            with open(name + ".json", "wb") as f:
                f.write(cu_json(cu_record(poc, idx, wc, z, d, qp, name)))
            n += 1
    return n
