"""Measures the int32 issue roof of the B200 with fhc_measure_int_issue and writes profiles/int_issue_peak.json.
Usage (gpurun): python tools/int_issue_peak.py > gpurun_out/int_issue_peak.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthevc_b200 as fh

runs = [fh.measure_int_issue(0) for _ in range(3)]
q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader"],
                   capture_output=True, text=True).stdout.strip()
best = max(r[0] for r in runs)
print(json.dumps({"int32_issue_gops": round(best, 1), "runs_gops": [round(r[0], 1) for r in runs], "launch_ms": [round(r[1], 3) for r in runs],
                  "method": "fhc_measure_int_issue: 148 SMs x 8 blocks x 256 threads, independent chains of IMAD / IADD3 / LOP3 (one third each, "
                            "inline PTX, SASS checked: 164 IMAD / 160 IADD3 / 168 LOP3 per unrolled body), best of 4 timed launches "
                            "(CUDA events), thread-level operations per second",
                  "nominal_gops": round(148 * 128 * 1.965, 1), "nvidia_smi": q}, indent=1))
