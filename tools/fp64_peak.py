#!/usr/bin/env python
"""Record the FP64 FMA peak this repo's roofline is quoted against (it is not in MEASURED_PEAKS.json):
cs_fp64_peak (register-resident DFMA chains, 8 independent per thread, every SM full) run back to back while NVML samples
SM clock, power and throttle reasons.   python tools/fp64_peak.py > profiles/r2_fp64_peak.json   (on the GPU box)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import clearsky_b200 as cs  # noqa: E402

ctx = cs.default_context()
for _ in range(3):
    ctx.fp64_peak(20000)
runs = []
with bench.ClockSampler(0, period=0.02) as clk:
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 3.0:
        runs.append(ctx.fp64_peak(20000) / 1e12)
import torch  # noqa: E402
p = torch.cuda.get_device_properties(0)
sm = p.multi_processor_count
c = clk.summary()
out = {"what": "FP64 FMA peak, TFLOP/s (2 flop per DFMA)", "kernel": "cs_fp64_peak / dfma_kernel (csrc/cs_api.cu)", "iters_per_launch": 20000,
       "launches": len(runs), "max": max(runs), "median": sorted(runs)[len(runs) // 2], "min": min(runs),
       "theoretical": sm * 64 * 2 * (c.get("sm_max_mhz") or 0) * 1e6 / 1e12, "sms": sm, "device": p.name, "clocks": c,
       "dmma_note": "mma.sync.m8n8k4.f64 alone reaches the same rate and shares the datapath with DFMA (tools/micro/dmma_dfma_mix.cu: "
                    "16 DFMA + 2 DMMA per iteration take the SUM of their separate times), so this is the FP64 roofline of the SM, "
                    "whichever instruction is used"}
print(json.dumps(out, indent=1))
