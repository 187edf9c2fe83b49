"""Per-function summary of an ncu report of fhc_ctu_kernel (stall samples, executed instructions) - the kernel is one
launch, so the function level is where the profile is readable.  Usage: python tools/ncu_functions.py <report.ncu-rep> <lib.so>"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    fn, addr2fn, size = None, {}, collections.Counter()
    for l in dis.split("\n"):
        m = re.match(r"^([^\s.][^\s]*):$", l)
        if m and not l.startswith(".L"):
            fn = m.group(1)
            continue
        m2 = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m2 and fn:
            size[fn] += 1
            if "fhc_ctu_kernel" in fn:
                addr2fn[int(m2.group(1), 16)] = fn
    rows = list(csv.reader(src.split("\n")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(collections.Counter)
    base = None
    keys = ("# Samples", "stall_no_inst", "stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "Instructions Executed", "Thread Instructions Executed")
    for r in rows[2:]:
        if len(r) < len(hdr) or not re.match("^(0x)?[0-9a-f]+$", r[ix["Address"]]):
            continue
        a = int(r[ix["Address"]], 16)
        if base is None:
            base = a
        f = addr2fn.get(a - base, "?")
        for k in keys:
            try:
                agg[f][k] += float(r[ix[k]] or 0)
            except ValueError:
                pass

    def short(f):
        n = f.split("$")[-1]
        n = re.sub(r"^_ZN3fhc\d+", "", n)
        return re.split(r"E(P|R|v|i)", n)[0][:26]

    tot = sum(v["# Samples"] for v in agg.values()) or 1
    toti = sum(v["Instructions Executed"] for v in agg.values()) or 1
    rr = list(csv.reader(raw.split("\n")))
    rv = dict(zip(rr[0], rr[2]))
    print("kernel %.1f ms, %.2f G warp instructions, %.1f active lanes per instruction, issue active %.1f %%, L2->GCC instruction sectors %.2f G, dram %.1f + %.1f MB" % (
        float(rv["gpu__time_duration.sum"]), toti / 1e9, float(rv["smsp__thread_inst_executed_per_inst_executed.ratio"]),
        float(rv["smsp__issue_active.avg.pct_of_peak_sustained_active"]), float(rv["gcc__xbar2gcc_sectors.sum"]) / 1e9,
        float(rv["dram__bytes_read.sum"]), float(rv["dram__bytes_write.sum"])))
    print("%-28s %6s %7s %7s %7s %7s %7s %8s %6s %6s" % ("function", "KB", "samp%", "noinst", "longsb", "barrier", "wait", "Ginstr", "inst%", "lanes"))
    for f, v in sorted(agg.items(), key=lambda x: -x[1]["# Samples"])[:24]:
        s = v["# Samples"] or 1
        ie = v["Instructions Executed"] or 1
        print("%-28s %6.1f %6.1f%% %6.0f%% %6.0f%% %6.0f%% %6.0f%% %8.2f %5.1f%% %6.1f" % (
            short(f), size[f] * 16 / 1024, 100 * s / tot, 100 * v["stall_no_inst"] / s, 100 * v["stall_long_sb"] / s, 100 * v["stall_barrier"] / s,
            100 * v["stall_wait"] / s, ie / 1e9, 100 * ie / toti, v["Thread Instructions Executed"] / ie))


if __name__ == "__main__":
    main()
