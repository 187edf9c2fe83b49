"""Stall-reason mix per function of an ncu report of fhc_ctu_kernel (warp-state samples of the source page).
Usage: python tools/ncu_stalls.py <report.ncu-rep> <lib.so>"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, lib = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
fn=None; a2f={}
for l in dis.split("\n"):
    m = re.match(r"^([^\s.][^\s]*):$", l)
    if m and not l.startswith(".L"): fn=m.group(1); continue
    m2 = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m2 and fn and "fhc_ctu_kernel" in fn: a2f[int(m2.group(1),16)] = fn
rows = list(csv.reader(src.split("\n"))); hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter); base=None
for r in rows[2:]:
    if len(r) < len(hdr) or not re.match("^(0x)?[0-9a-f]+$", r[ix["Address"]]): continue
    a = int(r[ix["Address"]],16)
    if base is None: base=a
    f = a2f.get(a-base,"?")
    n = f.split("$")[-1]; n = re.sub(r"^_ZN3fhc\d+","",n); n = re.split(r"E(P|R|v|i)", n)[0][:18]
    for k in keys:
        try: agg[n][k] += float(r[ix[k]] or 0)
        except ValueError: pass
tot = collections.Counter()
skip = ("compress_picture","role_loop","_Z14fhc_ctu_kernel")
print("%-20s %7s " % ("function","samples") + " ".join("%6s" % k[6:12] for k in keys))
for n,v in sorted(agg.items(), key=lambda x:-sum(x[1].values()))[:16]:
    s = sum(v.values()) or 1
    print("%-20s %7d " % (n, s) + " ".join("%5.1f%%" % (100*v[k]/s) for k in keys))
    if n not in skip:
        for k in keys: tot[k]+=v[k]
s = sum(tot.values()) or 1
print("%-20s %7d " % ("busy code total", s) + " ".join("%5.1f%%" % (100*tot[k]/s) for k in keys))
