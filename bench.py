#!/usr/bin/env python3
"""bench.py - 1080p all-intra mode-decision throughput of the B200 engine (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload = BASELINE configs[1]: encoder_intra_main.cfg, 1920x1080 8-bit 4:2:0, the 64-frame synthetic clip at QP 22 / 27 /
32 / 37 = 256 picture jobs per GPU.  A step = one pass of the hot path (TEncCu::xCompressCU and everything below it, for
all 510 CTUs of every picture) over those 256 pictures, through the C ABI with HOST buffers: fhc_submit_frames (page-
locked H2D of the 256 pictures), the persistent kernel (picture chains scheduled CTU row by CTU row over the thread blocks),
fhc_fetch_picture_async of every picture's CTU records (page-locked D2H).  `e2e` is the wall time of the steps, `value` the CUDA-event time of the
kernel inside the same steps (inputs resident in HBM when it starts); both max over ranks.  N > 1: one process per GPU
(torchrun), every rank encodes its own contiguous 64-frame group (configs[2]: 512 frames on 8 GPUs, weak scaling), no
collective on the data path and no NCCL - a gloo group provides the barrier and the max over ranks.
Steps are pipelined (step s+1 is submitted before the results of step s are fetched: two submits in flight, their kernels
overlap at the tail); `value` is then the CUDA-event time from the first kernel's start to the last kernel's end.  At
N = 1 three extra objects follow from one more step each after the timed region (not part of `value`): `deblock`
(fhc_deblock_pictures on the last step's pictures), `cnn_guided` (BASELINE configs[4]) and `wavefront` (--WaveFrontSynchro=1).

`--impl reference` times the UNMODIFIED reference encoder (oracle/_ref/TAppEncoder, HM 16.14 with the OpenCV dump
stubbed) on the host cores: one process per core, one 1080p picture each per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
W, H, CLIP = 1920, 1080, 64
QPS = [22, 27, 32, 37]
CTUS = 30 * 17
RING = 8                            # page-locked result slots the fetched pictures cycle through
# algorithmic bytes per CTU = 6144 source + 12288 reconstruction + 12288 levels + ~6 KB metadata
ALGO_BYTES_PER_CTU = 6144 + 12288 + 12288 + 6144
# SURVEY.md 8(d) W_ctu = 13 S_rmd + sum_size w_size S_tq,size + 30 bins, counted exactly on this workload's QP mix by
# tools/work_count.py on the host emulation of the kernel source (profiles/r02_work_counts.md)
INT_OPS_PER_CTU = 42.1e6


def load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, name)))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10)
                self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.5)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def run_ours(args, rank, world, local_rank):
    import torch
    import fasthevc_b200 as fh
    from fasthevc_b200.sharding import frame_group
    from fasthevc_b200.synth import synth_frames

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")                   # host-side barrier / max only: the data path has no collective

    def barrier():
        if dist is not None:
            dist.barrier()

    # weak scaling: every rank owns a 64-frame group of the (64 * world)-frame sequence (configs[2]: 512 frames on 8 GPUs)
    first, last = frame_group(args.frames * world, rank, world)
    n_frames = last - first
    jobs = [(f, q) for q in QPS for f in range(n_frames)]          # QP-major: the expensive pictures start first
    n_jobs = len(jobs)
    t_gen = time.perf_counter()
    Y, U, V = synth_frames(W, H, n_frames, first_frame=first)
    t_gen = time.perf_counter() - t_gen
    sps = [fh.slice_params_from_qp(q) for _, q in jobs]

    enc = fh.Encoder(W, H, 8, max_frames=n_jobs, device=local_rank)
    host_in = enc.pinned_array((n_jobs, enc.frame_samples), np.uint8)
    packed = enc.pack_frames(Y, U, V)
    for j, (f, _) in enumerate(jobs):
        host_in[j] = packed[f]
    del packed, Y, U, V
    host_out = enc.pinned_array((RING, enc.num_ctus), fh.CTU_RESULT_DTYPE)
    rec_bytes = int(host_out[0].nbytes)
    int_peak_gops, _ = fh.measure_int_issue(local_rank)

    first_result_s = []
    # Pipelined submits (fhc_configure FHC_OPT_PIPELINED_SUBMITS): step s+1 is submitted before the results of step s are
    # fetched, so its host->device copy and the head of its kernel overlap the tail of step s (two submits may be outstanding).
    enc.set_pipelined(not args.no_pipeline)
    poc_next = [0]

    def submit_step():
        first = poc_next[0]
        poc_next[0] += n_jobs
        enc.submit(host_in, sps, first_poc=first)
        return first, time.perf_counter()

    def fetch_step(first, t_sub):
        bits = 0
        for j in range(n_jobs):
            enc.fetch_picture_async(first + j, host_out[j % RING])
            if j % RING == RING - 1 or j == n_jobs - 1:
                enc.fetch_sync()
                if j < RING:
                    first_result_s.append(time.perf_counter() - t_sub)
                bits += int(host_out[:(j % RING) + 1]["totalBits"].sum(dtype=np.int64))
        return bits

    def run_steps(k):
        """k passes over the 256 pictures with host buffers; returns the sum of the CTU bit counts."""
        bits = 0
        pending = submit_step()
        for _ in range(1, k):
            if args.no_pipeline:
                bits += fetch_step(*pending)
                pending = submit_step()
            else:
                nxt = submit_step()
                bits += fetch_step(*pending)
                pending = nxt
        return bits + fetch_step(*pending)

    t_w = time.perf_counter()
    if args.warmup:
        run_steps(args.warmup)
    t_w = time.perf_counter() - t_w

    sampler = ClockSampler(local_rank)
    launches0 = int(enc.lib.fhc_launch_count())
    del first_result_s[:]
    sampler.start()
    barrier()
    enc.span_begin()
    t0 = time.perf_counter()
    check_bits = run_steps(args.steps)
    my_ms = enc.span_ms()                      # device time from the first kernel's start to the last kernel's end
    e2e_local = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    launches = int(enc.lib.fhc_launch_count()) - launches0
    in_flight = fh.pictures_in_flight(local_rank)
    guided = wavefront = deblock = None
    if rank == 0 and world == 1 and not args.no_deblock:
        deblock = deblock_leg(enc, n_jobs, poc_next[0] - n_jobs)
    if rank == 0 and world == 1 and not args.no_cnn:
        guided = cnn_guided_leg(enc, host_in, host_out, sps, n_jobs, local_rank, poc_next[0] - n_jobs, poc_next[0])
    if rank == 0 and world == 1 and not args.no_wpp:
        wavefront = wavefront_leg(fh, host_in, host_out, sps, n_jobs, local_rank)
    enc.close()

    per_rank = [[my_ms / args.steps, e2e_local / args.steps]]
    if dist is not None:
        t = torch.tensor([my_ms, e2e_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(per_rank[0], dtype=torch.float64))
        per_rank = [[round(float(v), 1) for v in g.tolist()] for g in gathered]
        my_ms, e2e_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        total_jobs = n_jobs * world
        ms_per_step = my_ms / args.steps
        print("bench: %d jobs/GPU x %d GPU(s), kernel %.1f ms/step, e2e %.1f ms/step (synthesis %.1f s, warm-up %.1f s)"
              % (n_jobs, world, ms_per_step, e2e_ms / args.steps, t_gen, t_w), file=sys.stderr, flush=True)
        fps = total_jobs / (ms_per_step / 1e3)
        e2e_fps = total_jobs / (e2e_ms / args.steps / 1e3)
        peaks = load_json("MEASURED_PEAKS.json")
        hbm_peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        ctus_per_launch = CTUS * n_jobs
        launch_s = ms_per_step / 1e3
        int_achieved = INT_OPS_PER_CTU * ctus_per_launch / launch_s / 1e12        # Tops/s per launch (= per rank)
        hbm_achieved = ALGO_BYTES_PER_CTU * ctus_per_launch / launch_s / 1e9
        # DRAM bytes of one launch from the committed ncu capture of this very command (profiles/r02_launches_bench.csv),
        # scaled by the CTUs of the launch when the job count differs
        prof = (load_json(os.path.join("profiles", "r02_kernel_counters.json")) or {}).get("bench_launch")
        traffic = prof and round((prof["dram_bytes_read"] + prof["dram_bytes_written"]) * ctus_per_launch / prof["ctus"])
        line = {
            "metric": "1080p all-intra encode fps (byte-identical CTU decisions)" if (W, H) == (1920, 1080) else "SELF-TEST %dx%d (not the metric)" % (W, H),
            "value": round(fps, 4), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64 (u8 samples, i16 levels)",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: encoder_intra_main.cfg %dx%d 8-bit, %d-frame synthetic clip per GPU x QP{22,27,32,37} = "
                                   "%d picture jobs per GPU per step (one kernel launch; picture chains are scheduled CTU row by CTU row "
                                   "over %d chain slots per GPU, most work left first, the most expensive pictures alone in their block); "
                                   "warm-up steps are full steps; %s" % (W, H, n_frames, n_jobs, in_flight,
                                                                   "one submit at a time" if args.no_pipeline else "steps are pipelined: step s+1 is submitted before the results of step s are fetched (two submits in flight, their kernels overlap at the tail)"),
                       "jobs_per_gpu": n_jobs, "ctus_per_picture": CTUS, "parallelism": "frame-group dp%d, no collective, no NCCL" % world,
                       "l2_policy": "inputs + results of a step (%.1f GB) exceed the 126 MB L2; no explicit flush" % (n_jobs * (3.1 + 14.1) / 1e3)},
            "ctu_per_s": round(fps * CTUS, 1),
            "first_results_s": round(float(first_result_s[0]), 3),      # first submit of the timed region -> first 8 pictures on the host (the scheduler lets pictures finish together)
            "pipelined_submits": not args.no_pipeline,
            "e2e": {"value": round(e2e_fps, 4), "unit": "frames/s", "h2d_bytes_per_step": int(host_in.nbytes) * world,
                    "d2h_bytes_per_step": rec_bytes * n_jobs * world, "ms_per_step": round(e2e_ms / args.steps, 2)},
            "per_rank_ms_per_step": {"kernel_e2e": per_rank},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "int_issue", "achieved": round(int_achieved, 4), "peak": round(int_peak_gops / 1e3, 3), "unit": "Tops/s",
                         "frac": int_achieved / (int_peak_gops / 1e3),
                         "traffic": traffic,
                         "peak_source": "measured live by fhc_measure_int_issue (IMAD/IADD3/LOP3 mix at full occupancy); committed run: profiles/int_issue_peak.json",
                         "work_per_ctu": INT_OPS_PER_CTU,
                         "hbm": {"achieved": round(hbm_achieved, 4), "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                                 "bytes_per_ctu": ALGO_BYTES_PER_CTU},
                         "note": "SURVEY.md 8(d): the path is integer-issue / latency bound, not HBM or tensor bound; achieved = W_ctu x "
                                 "CTUs per launch / launch time, W_ctu counted by tools/work_count.py; traffic = DRAM read + write of one launch "
                                 "(ncu), almost all of it L2 write-backs of the chains' scratch (12 GB/s), see DESIGN.md 3"},
            "check_total_bits": check_bits,
        }
        if guided is not None:
            line["cnn_guided"] = guided
        if wavefront is not None:
            line["wavefront"] = wavefront
        if deblock is not None:
            line["deblock"] = deblock
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_both_legs()
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def deblock_leg(enc, n_jobs, first_poc):
    """SURVEY 8f rank 1, first stage of the per-picture tail - NOT part of `value`: TComLoopFilter::loopFilterPic for the 256
    pictures of the last timed step, on the device from the CTU records (fhc_deblock_pictures).  HBM-bound byte work:
    algorithmic bytes per picture = the records' reconstruction read (2 B / sample) + depth / trIdx (512 B / CTU) + the
    deblocked picture written (1 B / sample)."""
    out = enc.pinned_array((n_jobs, enc.frame_samples), np.uint8)
    enc.deblock_pictures(first_poc, n_jobs, out)                  # warm-up (allocates the planes)
    t0 = time.perf_counter()
    _, ms = enc.deblock_pictures(first_poc, n_jobs, out)
    wall = time.perf_counter() - t0
    algo = n_jobs * (enc.frame_samples * 3 + CTUS * 512)
    peaks = load_json("MEASURED_PEAKS.json") or {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = algo / (ms / 1e3) / 1e9
    return {"value": round(n_jobs / (ms / 1e3), 1), "unit": "frames/s", "kernel_ms": round(ms, 3), "pictures": n_jobs,
            "e2e": {"value": round(n_jobs / wall, 1), "unit": "frames/s", "d2h_bytes": int(out.nbytes)},
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bytes_per_picture": algo // n_jobs,
                         "note": "four kernels (gather, vertical edges, horizontal edges, pack); the passes re-read and re-write the int16 planes, "
                                 "so DRAM traffic is about 3x the algorithmic bytes"},
            "checksum": int(out[::37, ::101].sum(dtype=np.int64))}


def wavefront_leg(fh, host_in, host_out, sps, n_jobs, device):
    """SURVEY 8f rank 4, one extra step after the timed region - NOT part of `value`: the same 256 jobs with
    --WaveFrontSynchro=1 (fhc_seq_params.waveFrontSynchro).  A different bitstream than the headline's (the reference run
    with that flag writes the same one, tests/test_wpp.py): the CTU rows of a picture are chains of their own, so no
    picture has to run alone in its block, and one picture alone finishes in a fraction of the time."""
    enc = fh.Encoder(W, H, 8, max_frames=n_jobs, device=device, wpp=True)
    try:
        t0 = time.perf_counter()
        enc.submit(host_in, sps)
        first = None
        for j in range(n_jobs):
            enc.fetch_picture_async(j, host_out[j % RING])
            if j % RING == RING - 1 or j == n_jobs - 1:
                enc.fetch_sync()
                if first is None:
                    first = time.perf_counter() - t0
        wall = time.perf_counter() - t0
        ms = enc.last_kernel_ms()
        # latency of one picture alone on the device (QP 22, the most expensive one)
        enc.submit(host_in[:1], sps[:1], first_poc=n_jobs)
        enc.fetch_picture_async(n_jobs, host_out[0])
        enc.fetch_sync()
        ms1 = enc.last_kernel_ms()
    finally:
        enc.close()
    return {"value": round(n_jobs / (ms / 1e3), 4), "unit": "frames/s", "ms_per_step": round(ms, 2), "steps": 1,
            "e2e": {"value": round(n_jobs / wall, 4), "unit": "frames/s"}, "first_results_s": round(first, 3),
            "single_picture_ms": round(ms1, 1),
            "note": "one step of the 256 jobs with --WaveFrontSynchro=1 (the reference's own option; another bitstream than the "
                    "headline's, byte-identical to the reference run with the flag); single_picture_ms = one QP-22 picture alone on the device"}


def cnn_guided_leg(enc, host_in, host_out, sps, n_jobs, device, resident_poc, free_poc):
    """BASELINE configs[4] ("C2 with CNN-guided split"), one extra step after the timed region - NOT part of `value`:
    the tcgen05 split predictor (libfasthevc_cnn.so) runs on the pictures resident in the engine's input buffer, its
    decisions become fhc_submit_frames_hinted's hints, and the same 256 jobs are searched with them.  Nothing trained
    exists: seeded random weights, the two logits centred on this input so that about half of the depth-1
    CUs are split; the decisions equal those of the reference carrying the hook patch (tests/test_hints.py), not HM's."""
    from fasthevc_b200 import cnn
    wts = cnn.seeded_weights()
    dptr, _ = enc.device_picture_ptr(resident_poc)          # the pictures of the last timed step are still on the device
    sp = cnn.SplitPredictor(wts, device)
    logits, _ = sp.predict_device(dptr, 1, 8, W, H, n_jobs)
    wts["fc_b"] = np.array([0.0, -float(np.median(logits[..., 1] - logits[..., 0]))], np.float32)
    sp.close()
    sp = cnn.SplitPredictor(wts, device)
    t0 = time.perf_counter()
    logits, split = sp.predict_device(dptr, 1, 8, W, H, n_jobs)
    cnn_wall = time.perf_counter() - t0
    cnn_ms = sp.rerun_resident(10)
    cus = int(split.size)
    sp.close()
    hints = cnn.hints_from_split(W, H, split)
    t0 = time.perf_counter()
    enc.submit(host_in, sps, first_poc=free_poc, hints=hints)
    for j in range(n_jobs):
        enc.fetch_picture_async(free_poc + j, host_out[j % RING])
        if j % RING == RING - 1 or j == n_jobs - 1:
            enc.fetch_sync()
    wall = time.perf_counter() - t0
    ms = enc.last_kernel_ms()
    peaks = load_json("MEASURED_PEAKS.json") or {}
    tens_peak = float(peaks.get("bf16_dense_tflops", peaks.get("bf16_tflops", 0)) or 0) or None
    tflops = 5.62e6 * cus / (cnn_ms / 1e3) / 1e12
    return {"value": round(n_jobs / (ms / 1e3), 4), "unit": "frames/s", "ms_per_step": round(ms, 2), "steps": 1,
            "e2e": {"value": round(n_jobs / (wall + cnn_wall), 4), "unit": "frames/s", "note": "hinted submit + fetch of every picture + the predictor call (logits D2H)"},
            "split_fraction": round(float(split.mean()), 4),
            "predictor": {"cus": cus, "kernel_ms": round(cnn_ms, 4), "cus_per_s": round(cus / (cnn_ms / 1e3), 1),
                          "roofline": {"bound": "tensor", "achieved": round(tflops, 3), "peak": tens_peak, "unit": "TFLOP/s",
                                       "frac": (tflops / tens_peak) if tens_peak else None, "flop_per_cu": 5.62e6,
                                       "note": "2.81 M MAC per 32x32 CU (SURVEY 8 a20); fp16 operands, fp32 accumulation in tensor memory; "
                                               "the whole network of a CU stays in shared / tensor memory, so the kernel is bound by its "
                                               "per-CU epilogues and barriers, not by the tensor pipe"}},
            "note": "one step of the 256 jobs with CNN hints, seeded weights, logits centred on this input; not the headline metric"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified HM encoder on the host cores
# ------------------------------------------------------------------------------------------------------------------
def ref_paths():
    d = os.path.join(ROOT, "oracle", "_ref")
    return os.path.join(d, "TAppEncoder"), os.path.join(d, "cfg", "encoder_intra_main.cfg")


def ref_clip(workdir):
    from fasthevc_b200.synth import synth_frames, write_yuv
    yuv = os.path.join(workdir, "clip.yuv")
    if not os.path.exists(yuv):
        Y, U, V = synth_frames(W, H, 4)
        write_yuv(yuv, Y, U, V)
    return yuv


def ref_cmd(workdir, yuv, frame, qp, tag):
    exe, cfg = ref_paths()
    return [exe, "-c", cfg, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-fs", str(frame), "-f", "1", "-q", str(qp),
            "-b", os.path.join(workdir, "o%s.bin" % tag), "-o", ""]


def cpu_reference_step(workdir, cores, step_idx):
    """One picture per core, all cores in parallel; returns (pictures, seconds)."""
    yuv = ref_clip(workdir)
    procs = []
    t0 = time.perf_counter()
    for c in range(cores):
        cmd = ref_cmd(workdir, yuv, c % 4, QPS[(c + step_idx) % 4], str(c))
        procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=workdir))
    for p in procs:
        p.wait()
    return cores, time.perf_counter() - t0


def cpu_reference_sample(steps, warmup):
    exe, _ = ref_paths()
    if not os.path.exists(exe):
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/TAppEncoder not built"}
    cores = max(1, len(os.sched_getaffinity(0)))
    d = tempfile.mkdtemp()
    for i in range(warmup):
        cpu_reference_step(d, cores, i)
    n, t = 0, 0.0
    for i in range(steps):
        a, b = cpu_reference_step(d, cores, warmup + i)
        n += a
        t += b
    return {"value": round(n / t, 4), "unit": "frames/s", "cores": cores, "kind": "reference", "ctu_per_s": round(n / t * CTUS, 1),
            "sample": "oracle/_ref/TAppEncoder (HM 16.14, g++ -O2 -msse4.1): %d step(s) of one %dx%d picture per core, QP 22/27/32/37 "
                      "cycled over cores, one process per core" % (steps, W, H),
            "ms_per_step": round(t / steps * 1e3, 1)}


def cpu_baseline_both_legs():
    """(A) single process on one core, (B) one process per host core; bounded samples."""
    exe, _ = ref_paths()
    if not os.path.exists(exe):
        return {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/TAppEncoder not built"}
    b = cpu_reference_sample(1, 0)
    d = tempfile.mkdtemp()
    yuv = ref_clip(d)
    cmd = ref_cmd(d, yuv, 0, 32, "single")
    core0 = sorted(os.sched_getaffinity(0))[0]
    t0 = time.perf_counter()
    subprocess.run(["taskset", "-c", str(core0)] + cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d)
    t1 = time.perf_counter() - t0
    b["single_process"] = {"value": round(1.0 / t1, 4), "unit": "frames/s", "cores": 1, "ctu_per_s": round(CTUS / t1, 1),
                           "sample": "one %dx%d picture, QP 32, taskset -c %d" % (W, H, core0)}
    return b


def run_reference(args, rank, world):
    if rank != 0:
        return
    exe, _ = ref_paths()
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/TAppEncoder missing (build with oracle/build_ref.py)"}))
        return
    cb = cpu_reference_sample(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "1080p all-intra encode fps (byte-identical CTU decisions)", "value": cb["value"],
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": "encoder_intra_main.cfg 1920x1080 8-bit synthetic, QP{22,27,32,37}; bounded sample: one picture "
                                   "per host core per step, whole TAppEncoder (the hot path is 98.8 % of it, SURVEY F10)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=CLIP, help="frames of the clip per GPU (64 = configs[1]; anything else is a self-test)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-deblock", action="store_true", help="skip the extra deblocking step (fhc_deblock_pictures) after the timed region")
    ap.add_argument("--no-pipeline", action="store_true", help="one submit at a time: a step ends before the next one is submitted")
    ap.add_argument("--no-wpp", action="store_true", help="skip the extra wavefront step (--WaveFrontSynchro=1) after the timed region")
    ap.add_argument("--no-cnn", action="store_true", help="skip the extra CNN-guided step (BASELINE configs[4]) after the timed region")
    ap.add_argument("--size", default="1920x1080", help="picture size; anything but the default is a script self-test, not the metric")
    args = ap.parse_args()
    global W, H, CTUS
    W, H = [int(v) for v in args.size.split("x")]
    CTUS = ((W + 63) // 64) * ((H + 63) // 64)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
