#!/usr/bin/env python3
"""Summarise an ncu report of fhc_compress_pictures_kernel into profiles/: key raw metrics, stall reasons and
stall samples per source function (needs nvdisasm line info of the same build: fasthevc_b200/libfasthevc_cuda.so).
usage: tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/r01_ncu_summary.md"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "launch__block_size", "launch__grid_size"]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = ["# ncu summary: %s" % os.path.basename(rep), "", "## raw metrics (ncu --page raw)", ""]
    rows = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            lines.append("* `%s` = %s %s" % (h, v, u))
    # source-level stall samples
    sass = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h2, data = sass[1], sass[2:]
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "fasthevc_b200", "libfasthevc_cuda.so")], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    seq, cur = [], (None, 0)
    for ln in run(["nvdisasm", "-g", "-c", os.path.join(d, cub)]).split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            seq.append(cur)
    src = open(os.path.join(ROOT, "fasthevc_b200", "csrc", "fhc_core.cuh")).read().split("\n")
    funcs = [(i, m.group(2)) for i, l in enumerate(src, 1) for m in [re.match(r"^(FHC_DH|FHC_D|FHC_NOINLINE|FHC_HD)\s+[\w\s\*]+?\b(\w+)\(", l)] if m]

    def fn(c):
        if c[0] != "fhc_core.cuh":
            return "%s (kernel entry / worker wait loop)" % c[0]
        name = "?"
        for i, n in funcs:
            if i <= c[1]:
                name = n
            else:
                break
        return name

    isamp, iex = h2.index("# Samples"), h2.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    fs, fe, st = collections.Counter(), collections.Counter(), collections.Counter()
    aligned = len(seq) == len(data)
    for k, r in enumerate(data):
        key = fn(seq[k]) if aligned else "?"
        fs[key] += int(r[isamp] or 0)
        fe[key] += int(r[iex] or 0)
        if aligned and seq[k][0] == "fhc_core.cuh":
            for i in stall_cols:
                if r[i]:
                    st[h2[i]] += int(float(r[i]))
    tot, te = sum(fs.values()), sum(fe.values())
    lines += ["", "## warp stall samples and executed instructions per source function (%s)" % ("line info aligned" if aligned else "NOT aligned"), "",
              "| function | samples % | instructions % |", "|---|---|---|"]
    for k, v in fs.most_common(22):
        lines.append("| %s | %.1f | %.1f |" % (k, 100.0 * v / tot, 100.0 * fe[k] / te))
    s = sum(st.values())
    lines += ["", "## stall reasons inside the decision/worker code (fhc_core.cuh only)", ""]
    for k, v in st.most_common(10):
        lines.append("* %s: %.1f %%" % (k, 100.0 * v / max(1, s)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
