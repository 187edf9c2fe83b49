"""Executed-instruction histogram by SASS opcode of an ncu report of fhc_ctu_kernel (source page), plus the static count
in the library.  Usage: python tools/ncu_opcodes.py <report.ncu-rep> <lib.so>"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.split("\n")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    dyn, thr, stat = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) or not re.match("^(0x)?[0-9a-f]+$", r[ix["Address"]]):
            continue
        sass = r[ix["Source"]].strip()
        m = re.match(r"^(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", sass)
        if not m:
            continue
        op = m.group(2)
        base = op.split(".")[0]
        key = base
        if base in ("LD", "ST", "LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "ATOMS", "RED"):
            key = base + ("." + ".".join(p for p in op.split(".")[1:] if p in ("E", "128", "64", "U8", "S8", "U16", "S16")) if "." in op else "")
        n = float(r[ix["Instructions Executed"]] or 0)
        dyn[key] += n
        thr[key] += float(r[ix["Thread Instructions Executed"]] or 0)
        stat[key] += 1
    tot = sum(dyn.values()) or 1
    print("| opcode | executed warp instructions | share | active lanes | static |")
    print("|---|---|---|---|---|")
    for k, v in dyn.most_common(45):
        print("| %s | %.3f G | %.2f %% | %.1f | %d |" % (k, v / 1e9, 100 * v / tot, thr[k] / max(v, 1), stat[k]))
    grp = lambda names: sum(dyn[k] for k in dyn if k.split(".")[0] in names)
    print()
    print("total %.2f G warp instructions, %.1f active lanes on average" % (tot / 1e9, sum(thr.values()) / tot))
    print("shared-memory loads/stores (LDS/STS) %.2f %%, generic/global (LD/ST/LDG/STG) %.2f %%, local (LDL/STL) %.2f %%, shuffles (SHFL) %.2f %%, fp64 (DADD/DMUL/DFMA/DSETP/I2F.F64...) %.2f %%"
          % (100 * grp(("LDS", "STS")) / tot, 100 * grp(("LD", "ST", "LDG", "STG")) / tot, 100 * grp(("LDL", "STL")) / tot, 100 * grp(("SHFL",)) / tot,
             100 * grp(("DADD", "DMUL", "DFMA", "DSETP", "I2F", "F2I")) / tot))


if __name__ == "__main__":
    main()
