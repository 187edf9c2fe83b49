#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: build the real reference (HM 16.14 as shipped in
omricarmi/FastHEVC) into oracle/_ref/ so it can pin the C restatement in
oracle/hm_intra.c and serve as the CPU baseline (`cpu_baseline.kind = "reference"`).

Nothing from /root/reference is copied into the repository history: the
sources are compiled where they lie, except for the handful of files that need
the HARP/OpenCV dataset-dump hooks stubbed; those are patched
into a scratch directory under oracle/_ref/ that is deleted after the build.

Outputs (all git-ignored, all under oracle/_ref/):
  TAppEncoder      stock encoder, HARP hooks removed
  TAppEncoderDump  same + oracle/ref_dump_hook.cpp (per-CTU golden dumps via FHC_DUMP=file)
  TAppDecoder      stock decoder (conformance checker: MD5 SEI)
  TAppEncoderStages stock encoder with HM's debug switches DEBUG_INTRA_SEARCH_COSTS / DEBUG_TRANSFORM_AND_QUANTISE on
                   (per-stage golden vectors, tests/golden/make_stage_golden.py)
  TAppEncoderDumpHinted  TAppEncoderDump + the CU-split hint hook (two conditions in TEncCu::xCompressCU, hints from $FHC_HINTS):
                   pins the oracle's hinted mode (hm_intra_compress_picture_hinted) against the reference's own arithmetic
  TAppEncoderFHC   the reference encoder with the drop-in boundary wired in: TEncCu::compressCtu calls the
                   TLibEncoderCUDA shim (fasthevc_b200/shim/TEncCuCUDA.cpp) instead of xCompressCU(depth 0) and is
                   linked against libfasthevc_cuda.so (RUNPATH $ORIGIN/../../fasthevc_b200, so LD_LIBRARY_PATH can
                   substitute the tests' emulated ABI) - whole-bitstream byte identity, end to end
  harp_json_ref    oracle/harp_json_ref.cpp + the reference's vendored jsoncpp: serialises one CU record of the dataset dumper
  cfg/             the two cfg files the metric uses (tiny text files, needed on the GPU box)

Patched sites (reference file:line):
  source/App/TAppEncoder/encmain.cpp:47-49,99-110   HARP includes + Global init + consts.json
  source/App/TAppEncoder/TAppEncCfg.cpp:49,1587     HARP include + Global.inputFilename
  source/Lib/TLibEncoder/TEncGOP.cpp:49,1562        HARP include + CShow_PredResiReco dump
  source/Lib/TLibEncoder/TEncCu.cpp:260             (TAppEncoderFHC only) xCompressCU(depth 0) -> fhc_shim_compress_ctu
  source/Lib/TLibEncoder/TEncGOP.cpp:1619           (TAppEncoderFHC only) m_pcLoopFilter->loopFilterPic -> fhc_shim_loop_filter
  source/Lib/TLibEncoder/TEncTop.cpp:57             (TAppEncoderFHC only) first POC from FHC_POC_BASE (frame groups on several host processes)
  source/Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp:253  (TAppEncoderFHC) getStatistics skipped when the shim has the device's statistics;
                                                    (TAppEncoderDump) fhc_dump_sao_stats after it
"""
import os, re, shutil, subprocess, sys, concurrent.futures as cf

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FHC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-std=c++11", "-O2", "-msse4.1", "-w"]


def patched_sources(scratch):
    """Return {orig_path: patched_path} for the files that need the HARP stub."""
    src = os.path.join(REF, "source")
    out = {}

    def put(rel, text, tag=""):
        # keep the directory layout so that quoted relative includes still resolve
        dst = os.path.join(scratch, "source" + tag, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write(text)
        return dst

    # encmain.cpp
    rel = "App/TAppEncoder/encmain.cpp"
    t = open(os.path.join(src, rel)).read()
    t = re.sub(r'#include "\.\./\.\./\.\./Src_HARP/[^"]+"\n', "", t)
    a = t.index("    //GLOBAL INIT")
    b = t.index("    // starting time")
    t = t[:a] + t[b:]
    out[rel] = put(rel, t)

    rel = "App/TAppEncoder/TAppEncCfg.cpp"
    t = open(os.path.join(src, rel)).read()
    t = re.sub(r'#include "\.\./\.\./\.\./Src_HARP/[^"]+"\n', "", t)
    t = t.replace("Global.inputFilename = match[1];", "")
    out[rel] = put(rel, t)

    rel = "Lib/TLibEncoder/TEncGOP.cpp"
    t = open(os.path.join(src, rel)).read()
    t = re.sub(r'#include "\.\./\.\./\.\./Src_HARP/[^"]+"\n', "", t)
    t = t.replace("CShow_PredResiReco Show_PredResiReco(pcPic);", "")
    out[rel] = put(rel, t)
    return out


def dump_sources(scratch):
    """Instrumented variants: one call per finished CTU / per slice into ref_dump_hook.cpp."""
    src = os.path.join(REF, "source")
    out = {}
    rel = "Lib/TLibEncoder/TEncCu.cpp"
    t = open(os.path.join(src, rel)).read()
    t = t.replace("Void TEncCu::compressCtu( TComDataCU* pCtu )\n{",
                  "void fhc_dump_ctu(TComDataCU*);\nVoid TEncCu::compressCtu( TComDataCU* pCtu )\n{\n  struct FhcAtExit { TComDataCU* c; ~FhcAtExit() { fhc_dump_ctu(c); } } fhcAtExit = { pCtu };", 1)
    dst = os.path.join(scratch, "dump", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst

    rel = "Lib/TLibEncoder/TEncSlice.cpp"
    t = open(os.path.join(src, rel)).read()
    marker = "  for( UInt ctuTsAddr = startCtuTsAddr; ctuTsAddr < boundingCtuTsAddr; ++ctuTsAddr )\n  {\n    const UInt ctuRsAddr = pcPic->getPicSym()->getCtuTsToRsAddrMap(ctuTsAddr);\n    // initialize CTU encoder"
    assert t.count(marker) == 1, t.count(marker)
    t = t.replace(marker, "  fhc_dump_slice(pcPic, pcSlice, m_pcRdCost, m_pcTrQuant);\n" + marker)
    t = t.replace("Void TEncSlice::compressSlice(", "void fhc_dump_slice(TComPic*, TComSlice*, TComRdCost*, TComTrQuant*);\nVoid TEncSlice::compressSlice(", 1)
    dst = os.path.join(scratch, "dump", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst

    # SAO statistics of every picture (pins oracle/hm_sao.c): one call after getStatistics in SAOProcess
    rel = "Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp"
    t = open(os.path.join(src, rel)).read()
    call = "  getStatistics(m_statData, orgYuv, srcYuv, pPic);\n"
    assert t.count(call) == 1, t.count(call)
    t = t.replace(call, call + "  fhc_dump_sao_stats((void*)m_statData, m_numCTUsPic);\n")
    t = t.replace("Void TEncSampleAdaptiveOffset::SAOProcess(", "void fhc_dump_sao_stats(void*, int);\nVoid TEncSampleAdaptiveOffset::SAOProcess(", 1)
    dst = os.path.join(scratch, "dump", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst
    return out


def hinted_sources(scratch, dump):
    """TAppEncoderDumpHinted: the dump variant plus the CU-split hint hook exactly as INTEGRATION.md states it - two
    conditions in TEncCu::xCompressCU (TEncCu.cpp:576: the CU's own modes are skipped on 'split'; :892: the sub-branch
    is skipped on 'do not split'), for depth-1 CUs inside the picture; hints come from the file $FHC_HINTS."""
    rel = "Lib/TLibEncoder/TEncCu.cpp"
    t = open(dump[rel]).read()
    t = t.replace("void fhc_dump_ctu(TComDataCU*);", "void fhc_dump_ctu(TComDataCU*);\nint fhc_hint(TComDataCU*);", 1)
    a = "    if ( !bBoundary )\n    {\n        for (Int iQP=iMinQP; iQP<=iMaxQP; iQP++)"
    assert t.count(a) == 1, t.count(a)
    t = t.replace(a, "    const int fhcHint = (uiDepth == 1 && !bBoundary) ? fhc_hint(rpcBestCU) : 0;\n" + a.replace("!bBoundary", "!bBoundary && fhcHint != 2"))
    b = "const Bool bSubBranch = bBoundary || "
    assert t.count(b) == 1, t.count(b)
    t = t.replace(b, "const Bool bSubBranch = (fhcHint != 1) && ")
    c = "rpcBestCU->isSkipped(0) );"
    i = t.index("const Bool bSubBranch")
    j = t.index(c, i)
    t = t[:i] + t[i:j].replace("(fhcHint != 1) && ", "(fhcHint != 1) && ( bBoundary || ") + "rpcBestCU->isSkipped(0) ) );" + t[j + len(c):]
    dst = os.path.join(scratch, "hinted", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    return {rel: dst}


def stages_sources(scratch):
    """TAppEncoderStages: the stock encoder with HM's own debug switches DEBUG_INTRA_SEARCH_COSTS and
    DEBUG_TRANSFORM_AND_QUANTISE (TypeDef.h:58-59) turned on in the two files that use them: it prints every first-pass
    mode cost and every TU on its way through transform and quantiser (tests/golden/make_stage_golden.py)."""
    src = os.path.join(REF, "source")
    out = {}
    for rel, macro in (("Lib/TLibEncoder/TEncSearch.cpp", "DEBUG_INTRA_SEARCH_COSTS"), ("Lib/TLibCommon/TComTrQuant.cpp", "DEBUG_TRANSFORM_AND_QUANTISE")):
        t = open(os.path.join(src, rel)).read()
        marker = "//! \\ingroup"
        assert t.count(marker) >= 1, rel
        t = t.replace(marker, "#undef %s\n#define %s 1\n#include <iostream>\n%s" % (macro, macro, marker), 1)
        dst = os.path.join(scratch, "stages", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write(t)
        out[rel] = dst
    return out


def fhc_sources(scratch, patched):
    """TEncCu.cpp of the integrated encoder: the single call site of the hot path goes to the shim."""
    src = os.path.join(REF, "source")
    rel = "Lib/TLibEncoder/TEncCu.cpp"
    t = open(os.path.join(src, rel)).read()
    call = "xCompressCU( m_ppcBestCU[0], m_ppcTempCU[0], 0 DEBUG_STRING_PASS_INTO(sDebug) );"
    assert t.count(call) == 1, t.count(call)
    t = t.replace(call, "fhc_shim_compress_ctu(this, pCtu);")
    t = t.replace("Void TEncCu::compressCtu( TComDataCU* pCtu )\n{", "void fhc_shim_compress_ctu(TEncCu*, TComDataCU*);\nVoid TEncCu::compressCtu( TComDataCU* pCtu )\n{", 1)
    dst = os.path.join(scratch, "fhc", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out = {rel: dst}
    # first stage of the per-picture tail on the device: the one call of the deblocking filter in
    # TEncGOP::compressGOP (TEncGOP.cpp:1619) goes to the shim (fhc_deblock_pictures; FHC_SHIM_DEBLOCK=0 keeps the host filter)
    rel = "Lib/TLibEncoder/TEncGOP.cpp"
    t = open(patched[rel]).read()
    call = "    m_pcLoopFilter->loopFilterPic( pcPic );\n\n    /////"
    assert t.count(call) == 1, t.count(call)
    t = t.replace(call, "    fhc_shim_loop_filter( m_pcLoopFilter, pcPic );\n\n    /////")
    t = t.replace("Void TEncGOP::compressGOP(", "void fhc_shim_loop_filter(TComLoopFilter*, TComPic*);\nVoid TEncGOP::compressGOP(", 1)
    # (frame groups, see TEncTop.cpp below: the end-of-sequence test counts pictures of this instance, not absolute POCs)
    eos = "    if(pocCurr>=m_pcCfg->getFramesToBeEncoded())\n"
    assert t.count(eos) == 1, t.count(eos)
    t = t.replace(eos, "    if(pocCurr - (getenv(\"FHC_POC_BASE\") ? atoi(getenv(\"FHC_POC_BASE\")) : 0) >= m_pcCfg->getFramesToBeEncoded())\n")
    dst = os.path.join(scratch, "fhc", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst
    # second stage: the SAO statistics come from the device with the deblocked picture; the one getStatistics call of
    # SAOProcess (TEncSampleAdaptiveOffset.cpp:253) is skipped when the shim has them
    rel = "Lib/TLibEncoder/TEncSampleAdaptiveOffset.cpp"
    t = open(os.path.join(src, rel)).read()
    call = "  getStatistics(m_statData, orgYuv, srcYuv, pPic);\n"
    assert t.count(call) == 1, t.count(call)
    t = t.replace(call, "  if (!fhc_shim_sao_statistics((void*)m_statData, pPic)) getStatistics(m_statData, orgYuv, srcYuv, pPic);\n")
    t = t.replace("Void TEncSampleAdaptiveOffset::SAOProcess(", "bool fhc_shim_sao_statistics(void*, TComPic*);\nVoid TEncSampleAdaptiveOffset::SAOProcess(", 1)
    dst = os.path.join(scratch, "fhc", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst
    # frame groups on several host processes: an encoder instance that starts in the middle of the sequence
    # (--FrameSkip k) numbers its pictures from POC k instead of 0 when FHC_POC_BASE=k is set (TEncTop.cpp:57); with the
    # all-intra cfg every access unit is self-contained (CRA + ReWriteParamSetsFlag), so the groups' bitstreams concatenate
    # to the single-process bitstream
    # (the per-frame delta-QP table is indexed by absolute POC: TAppEncCfg.cpp:1404 sizes it for the group's first POC too)
    rel = "App/TAppEncoder/TAppEncCfg.cpp"
    t = open(patched[rel]).read()
    a1 = "m_aidQP = new Int[ m_framesToBeEncoded + m_iGOPSize + 1 ];"
    a2 = "::memset( m_aidQP, 0, sizeof(Int)*( m_framesToBeEncoded + m_iGOPSize + 1 ) );"
    assert t.count(a1) == 1 and t.count(a2) == 1
    base = "(getenv(\"FHC_POC_BASE\") ? atoi(getenv(\"FHC_POC_BASE\")) : 0)"
    t = t.replace(a1, "m_aidQP = new Int[ m_framesToBeEncoded + m_iGOPSize + 1 + %s ];" % base)
    t = t.replace(a2, "::memset( m_aidQP, 0, sizeof(Int)*( m_framesToBeEncoded + m_iGOPSize + 1 + %s ) );" % base)
    dst = os.path.join(scratch, "fhc", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst
    rel = "Lib/TLibEncoder/TEncTop.cpp"
    t = open(os.path.join(src, rel)).read()
    init = "  m_iPOCLast          = -1;\n"
    assert t.count(init) == 1, t.count(init)
    t = t.replace(init, "  m_iPOCLast          = (getenv(\"FHC_POC_BASE\") ? atoi(getenv(\"FHC_POC_BASE\")) : 0) - 1;\n")
    dst = os.path.join(scratch, "fhc", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(t)
    out[rel] = dst
    return out


def compile_one(job):
    srcfile, obj, incs, extra = job
    if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(srcfile):
        return obj
    cc = "gcc" if srcfile.endswith(".c") else CXX
    flags = ["-O2", "-w"] if srcfile.endswith(".c") else CXXFLAGS
    cmd = [cc] + flags + extra + incs + ["-c", srcfile, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return obj


def main():
    if not os.path.isdir(os.path.join(REF, "source")):
        print("build_ref: %s not present; keeping prebuilt oracle/_ref" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    scratch = os.path.join(OUT, "scratch")
    objdir = os.path.join(OUT, "obj")
    shutil.rmtree(scratch, ignore_errors=True)
    os.makedirs(scratch); os.makedirs(objdir, exist_ok=True)
    src = os.path.join(REF, "source")
    stub = patched_sources(scratch)
    dump = dump_sources(scratch)
    fhc = fhc_sources(scratch, stub)
    stages = stages_sources(scratch)
    hinted = hinted_sources(scratch, dump)
    incs = ["-I" + os.path.join(src, "Lib"), "-I" + os.path.join(src, "App"), "-I" + os.path.join(src, "App/TAppEncoder")]

    def listing(d, exts=(".cpp", ".c")):
        return sorted(os.path.join(d, f) for f in os.listdir(os.path.join(src, d)) if f.endswith(exts))

    common = listing("Lib/TLibCommon") + listing("Lib/TLibVideoIO") + listing("Lib/TAppCommon") + listing("Lib/libmd5")
    enc = listing("Lib/TLibEncoder") + listing("App/TAppEncoder")
    dec = listing("Lib/TLibDecoder") + listing("App/TAppDecoder")

    jobs, objs = [], {"enc": [], "dump": [], "dec": [], "fhc": [], "stages": [], "hinted": []}
    def add(rel, which):
        tagged = rel.replace("/", "_")
        if which == "fhc" and rel in fhc:
            s, o = fhc[rel], os.path.join(objdir, "fhc_" + tagged + ".o")
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        elif which == "hinted" and rel in hinted:
            s, o = hinted[rel], os.path.join(objdir, "hinted_" + tagged + ".o")
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        elif which == "hinted" and rel in dump:
            s, o = dump[rel], os.path.join(objdir, "dump_" + tagged + ".o")
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        elif which == "stages" and rel in stages:
            s, o = stages[rel], os.path.join(objdir, "stages_" + tagged + ".o")
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        elif which == "dump" and rel in dump:
            s, o = dump[rel], os.path.join(objdir, "dump_" + tagged + ".o")
            # quoted relative includes resolve against the original directory
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        elif rel in stub:
            s, o = stub[rel], os.path.join(objdir, "stub_" + tagged + ".o")
            extra = ["-iquote", os.path.dirname(os.path.join(src, rel))]
        else:
            s, o = os.path.join(src, rel), os.path.join(objdir, tagged + ".o")
            extra = []
        if (s, o, incs, extra) not in jobs and not any(j[1] == o for j in jobs):
            jobs.append((s, o, incs, extra))
        objs[which].append(o)

    for rel in common:
        add(rel, "enc"); add(rel, "dump"); add(rel, "dec"); add(rel, "fhc"); add(rel, "stages"); add(rel, "hinted")
    for rel in enc:
        add(rel, "enc"); add(rel, "dump"); add(rel, "fhc"); add(rel, "stages"); add(rel, "hinted")
    for rel in dec:
        add(rel, "dec")
    hook_o = os.path.join(objdir, "ref_dump_hook.o")
    jobs.append((os.path.join(HERE, "ref_dump_hook.cpp"), hook_o, incs, []))
    objs["dump"].append(hook_o)
    objs["hinted"].append(hook_o)
    repo = os.path.dirname(HERE)
    shim_o = os.path.join(objdir, "TEncCuCUDA.o")
    jobs.append((os.path.join(repo, "fasthevc_b200", "shim", "TEncCuCUDA.cpp"), shim_o, incs + ["-I" + os.path.join(repo, "include")], []))
    objs["fhc"].append(shim_o)

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(compile_one, jobs))

    for name, key in (("TAppEncoder", "enc"), ("TAppEncoderDump", "dump"), ("TAppDecoder", "dec"), ("TAppEncoderStages", "stages"),
                      ("TAppEncoderDumpHinted", "hinted")):
        subprocess.check_call([CXX, "-O2", "-o", os.path.join(OUT, name)] + sorted(set(objs[key])) + ["-lm"])
    os.makedirs(os.path.join(OUT, "cfg"), exist_ok=True)
    for c in ("encoder_intra_main.cfg", "encoder_intra_main10.cfg"):
        shutil.copy(os.path.join(REF, "cfg", c), os.path.join(OUT, "cfg", c))
    libdir = os.path.join(repo, "fasthevc_b200")
    if os.path.exists(os.path.join(libdir, "libfasthevc_cuda.so")):
        subprocess.check_call([CXX, "-O2", "-o", os.path.join(OUT, "TAppEncoderFHC")] + sorted(set(objs["fhc"])) +
                              ["-L" + libdir, "-lfasthevc_cuda", "-Wl,--enable-new-dtags", "-Wl,-rpath,$ORIGIN/../../fasthevc_b200", "-lm"])
    else:
        print("build_ref: libfasthevc_cuda.so not built yet - TAppEncoderFHC skipped")
    # the reference's vendored jsoncpp (compiled where it lies) + oracle/harp_json_ref.cpp: the serialiser of the dataset
    # dumper's CU records, used to generate tests/golden/harp_cu_json.json
    subprocess.check_call([CXX, "-std=c++11", "-O1", "-w", "-I" + os.path.join(REF, "Src_HARP"), "-o", os.path.join(OUT, "harp_json_ref"),
                           os.path.join(HERE, "harp_json_ref.cpp"), os.path.join(REF, "Src_HARP", "jsoncpp.cpp")])
    shutil.rmtree(scratch, ignore_errors=True)
    shutil.rmtree(objdir, ignore_errors=True)
    # provenance of the binaries that travel to the GPU box (they are git-ignored): sha256 next to them
    import hashlib
    with open(os.path.join(OUT, "SHA256SUMS"), "w") as f:
        for name in sorted(os.listdir(OUT)):
            path = os.path.join(OUT, name)
            if os.path.isfile(path) and name != "SHA256SUMS":
                f.write("%s  %s\n" % (hashlib.sha256(open(path, "rb").read()).hexdigest(), name))
    print("build_ref: built", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
