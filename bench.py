#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 line-by-line radiative-transfer engine.

Metric (BASELINE.json): line×ν×layer evaluations per second on configs[1] --
"Earth-like clear sky: synthetic CO2+H2O ~500k Voigt lines, 0-3000 cm^-1 at 0.01 cm^-1, 100 layers, OLR on 1 B200"
(SURVEY.md section 8d, C2).  One step = one full pass of the hot path: Voigt line summation of both gases at the
101 levels -> layer optical depths -> 5-stream Schwarzschild sweeps -> spectral reduction to OLR / F+ / F-.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2small|c1]

N > 1 (launched by torchrun, one rank per GPU): contiguous ν slices balanced by evaluation count, lines within
slice ± cutoff only, and ONE NCCL all-reduce of the 2*np spectrally integrated fluxes per step.
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of
the same workload; the Julia reference itself cannot run in this image (no Julia runtime).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)

METRIC = "line×ν×layer evals/s"
UNIT = "evals/s"
FLOP_PER_EVAL = 10.0   # Voigt, far-wing dominated
# dram__bytes_read.sum + dram__bytes_write.sum of the line sum of one gas (101 levels) on C2 = its two launches,
# line_sum_kernel<VOIGT, COLD> (1.701 + 0.233 GB) + far_fold_kernel<4, 32> (1.066 + 0.235 GB), from the ncu --set full capture
# summarised in profiles/r2_ncu_full_line_sum_voigt_split.csv
NCU_TRAFFIC_BYTES_PER_LAUNCH = 1.701310e9 + 0.232902e9 + 1.065896e9 + 0.234923e9


# ------------------------------------------------------------------------------------------------
# workload
def synthetic_lines(cs, n, seed, M, γs_rng, νmax=3000.0):
    """SURVEY.md section 8(d) C2 generator: HITRAN .par field widths/rounding, numpy PCG64"""
    rng = np.random.default_rng(np.random.PCG64(seed))
    ν = np.sort(np.round(rng.uniform(0.0, νmax, n), 6))
    ν = np.maximum(ν, 1e-6)
    S = np.array([float(f"{x:.3E}") for x in 10.0 ** rng.uniform(-30, -19, n)])
    γa = np.round(rng.uniform(0.05, 0.10, n), 4)
    γs = np.round(rng.uniform(*γs_rng, n), 3)
    Epp = np.round(rng.uniform(0, 3000, n), 4)
    na = np.round(rng.uniform(0.5, 0.8, n), 2)
    mp = cs.MOLPARAM[M]
    return cs.SpectralLines(mp.name, mp.formula, n, M, np.ones(n, dtype=np.int16), np.full(n, mp.mu[0]),
                            np.full(n, mp.A[0]), ν, S, γa, γs, Epp, na)


def make_workload(cs, name):
    if name == "c2":
        nlines, nν, nlayer = 250_000, 300_000, 100
    elif name == "c2small":
        nlines, nν, nlayer = 25_000, 30_000, 20
    else:
        raise ValueError(name)
    co2 = synthetic_lines(cs, nlines, 20261018, 2, (0.06, 0.13))
    h2o = synthetic_lines(cs, nlines, 20261019, 1, (0.10, 0.50))
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    P = cs.pressuregrid(10.0, 1e5, nlayer + 1)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    T = Γ(P)
    return dict(name=name, gases=[(co2, 400e-6), (h2o, 1e-3)], ν=ν, P=P, T=T, μ=0.029, g=9.8, cut=25.0,
                nstream=5, nlob=2, nlayer=nlayer)


def balanced_slices(cost, n):
    from clearsky_b200 import sharding
    return sharding.balanced_slices(cost, n)


def per_point_counts(ν, νl, cut):
    from clearsky_b200 import sharding
    return sharding.per_point_counts(ν, νl, cut)


def slice_cost(ν, gases, cut):
    """cost model used to balance the ν slices (clearsky_b200/sharding.py)"""
    from clearsky_b200 import sharding
    return sharding.slice_cost(ν, [sl.ν for sl, _ in gases], cut)


def trapz_weights(ν):
    from clearsky_b200 import sharding
    return sharding.trapz_weights(ν)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML in-process: a few microseconds per
    sample, no nvidia-smi process spawns competing with the launches; nvidia-smi is the fallback)."""

    def __init__(self, device, period=0.1):
        self.device, self.period = device, period
        self.rows = []          # (sm_mhz, sm_max_mhz, power_w, reasons bitmask)
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and vis.split(",")[device].isdigit() else device
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample(self):
        nv = self._nvml
        if nv is not None:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.rows.append((float(sm), float(mx), pw, int(rs)))
                return
            except Exception:
                pass
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            bits = 0
            for i, b in enumerate((0x8, 0x40, 0x20, 0x4)):
                if out[3 + i].strip().lower().startswith("active"):
                    bits |= b
            self.rows.append((float(out[0]), float(out[1]), float(out[2]), bits))
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.period if self._nvml is not None else 1.0)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        bits = 0
        for r in self.rows:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": max(r[1] for r in self.rows),
                "power_w_max": max(r[2] for r in self.rows), "reasons": [n for b, n in names.items() if bits & b],
                "samples": len(self.rows), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
def host_threads():
    """host cores this process may use (torchrun exports OMP_NUM_THREADS=1 for nproc > 1: never rely on the OpenMP default)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(wl, world, total_evals):
    """the `config` object of the JSON line -- identical for the repo arm and the reference arm"""
    return {"workload": wl["name"], "lines": int(sum(len(sl.ν) for sl, _ in wl["gases"])), "n_nu": len(wl["ν"]),
            "layers": wl["nlayer"], "shape": "voigt", "cutoff_cm-1": wl["cut"], "nstream": wl["nstream"],
            "nlobatto": wl["nlob"], "evals_per_step": int(total_evals), "parallelism": f"nu-slices x{world}",
            "l2": "per-step working set (line records 3.2 GB + sigma 0.24 GB) exceeds the 126 MB L2"}


def total_evals_host(orc, wl):
    """exact iterations of surf!'s inner loop (line_shapes.jl:75-82) over the whole workload, counted on the host"""
    ν, cut = wl["ν"], wl["cut"]
    return sum(orc.count_evals(ν, orc.included_lines(ν, sl.ν, cut), cut) for sl, _ in wl["gases"]) * len(wl["P"])


def cpu_sample(cs, orc, wl, nthreads, target_evals=2.0e9):
    """bounded CPU sample of the same workload: a contiguous ν slice from the middle of the grid × all levels,
    both gases, oracle restatement parallel over levels (like bake's @threads, gases.jl:115) then fluxes.
    run() -> (Σ[nlev, n], flux dict): the values are kept, the GPU is checked against them (parity_sample)."""
    ν, P, T = wl["ν"], wl["P"], wl["T"]
    nlev = len(P)
    per_ν = sum(len(sl.ν) for sl, _ in wl["gases"]) * (2 * wl["cut"]) / 3000.0 * nlev
    n = int(min(len(ν), max(64, target_evals / per_ν)))
    i0 = (len(ν) - n) // 2
    νs = np.ascontiguousarray(ν[i0:i0 + n])
    evals = sum(orc.count_evals(νs, orc.included_lines(νs, sl.ν, wl["cut"]), wl["cut"]) for sl, _ in wl["gases"]) * nlev
    m, W = cs.streamnodes(wl["nstream"])
    x, w = cs.lobattonodes(wl["nlob"])
    μn = np.full((nlev - 1, wl["nlob"]), wl["μ"])

    def run():
        σ = np.zeros((nlev, n))
        for sl, C in wl["gases"]:
            σ += C * orc.xsec(orc.VOIGT, sl, νs, T, P, C * P, wl["cut"], nthreads=nthreads)
        F = orc.fluxes(νs, P, wl["nlob"], w, μn, T, σ, wl["g"], None, None, 0.841, wl["nstream"], m, W,
                       nthreads=nthreads, full=False)
        return σ, F

    sample = f"ν slice of {n} points (indices {i0}..{i0 + n - 1}) × all {nlev} levels, both gases, Voigt + fluxes"
    return run, evals, sample, (i0, n)


def run_reference(args):
    """the reference's own CPU algorithm for the path (oracle/ port: no Julia runtime in the image), all host threads,
    each step one bounded sample of the workload sized so that warm-up + steps end within a few minutes"""
    import clearsky_b200 as cs   # host-side generators/readers only: no CUDA call on this arm
    from oracle import oracle as orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc.build()
    wl = make_workload(cs, args.workload)
    nthreads = host_threads()
    if args.cpu_evals > 0:
        target = args.cpu_evals
    else:
        # calibrate on a small sample, then size the step so that (warm-up + steps) x t stays near 150 s
        run, evals, _, _ = cpu_sample(cs, orc, wl, nthreads, target_evals=1.0e9)
        run()
        t0 = time.perf_counter()
        run()
        rate = evals / (time.perf_counter() - t0)
        target = min(4.0e10, max(5.0e8, rate * 150.0 / max(1, args.steps + args.warmup)))
    run, evals, sample, _ = cpu_sample(cs, orc, wl, nthreads, target_evals=target)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    v = evals / dt
    cfg = workload_config(wl, args.gpus, total_evals_host(orc, wl))
    cfg["sample"] = sample
    cfg["note"] = ("CPU restatement of the reference (oracle/, C + OpenMP over levels like bake's @threads); the Julia "
                   "reference cannot run here (no Julia runtime in the image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line, ensure_ascii=False))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import clearsky_b200 as cs
    from clearsky_b200 import sharding
    from clearsky_b200._lib import check, f64, lib, ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints a version banner on stdout when its first communicator comes up; stdout must carry exactly one
        # JSON line, so file descriptor 1 points at stderr while the communicator is created (first collective included)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=f"cuda:{local}")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    # run the library on torch's current (non-default) stream so that torch.cuda.Event brackets its kernels
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx = cs.Context(local, stream=stream.cuda_stream)
    wl = make_workload(cs, args.workload)
    ν, P, T = wl["ν"], wl["P"], wl["T"]
    nlev = len(P)
    cut = wl["cut"]

    # ---- ν sharding: contiguous slices balanced by evaluations, global trapezoid weights
    edges = balanced_slices(slice_cost(ν, wl["gases"], cut), world)
    i0, i1 = edges[rank], edges[rank + 1]
    νs = np.ascontiguousarray(ν[i0:i1])
    wts = np.ascontiguousarray(trapz_weights(ν)[i0:i1])
    # every line the slice can see: the inclusive per-point rule decides inside the kernel; the strict prefilter of
    # includedlines is applied to the GLOBAL grid (cs_lines_set_grid_range), as in the unsharded run
    my_gases = [(sharding.slice_lines(sl, νs[0], νs[-1], cut, grid=(ν[0], ν[-1])), C) for sl, C in wl["gases"]]
    m, W = cs.streamnodes(wl["nstream"])
    x, w = cs.lobattonodes(wl["nlob"])
    m, W, w = f64(m), f64(W), f64(w)
    μn = f64(np.full((nlev - 1, wl["nlob"]), wl["μ"]))
    Tn, Pn = f64(T), f64(P)
    dF = torch.zeros(2 * nlev, dtype=torch.float64, device=f"cuda:{local}")

    # resident objects for the device-timed loop
    dls = [cs.DeviceLines(sl, ctx) for sl, _ in my_gases]
    # exact evaluation counts (iterations of surf!'s inner loop, line_shapes.jl:75-82) from the library's host counter
    my_evals = sum(dl.count_evals(νs, cut) for dl in dls) * nlev
    if world > 1:
        t_ev = torch.tensor([float(my_evals)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t_ev)
        total_evals = int(round(t_ev.item()))
    else:
        total_evals = my_evals
    ws = cs.SigmaWorkspace(νs, nlev, ctx)
    Cs = [f64(np.full(nlev, C)) for _, C in my_gases]

    def timed_step():
        """one pass of the hot path with inputs resident in HBM; no host synchronisation anywhere in it"""
        ws.zero()
        for dl, C in zip(dls, Cs):
            check(lib().cs_sigma_add_lines(ws.h, dl.h, cs._lib.CS_VOIGT, ptr(Tn), ptr(Pn), ptr(C), cut))
        check(lib().cs_fluxes_device(ws.h, nlev, ptr(Pn), wl["nlob"], ptr(w), ptr(μn), ptr(Tn), wl["g"], None, None,
                                     0.841, wl["nstream"], ptr(m), ptr(W), ptr(wts), dF.data_ptr()))
        if world > 1:
            dist.all_reduce(dF)
        return dF

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- end-to-end inputs: pinned host copies of everything a step uploads (the byte counts below are taken from
    # these very arrays)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()

    def pinned_lines(sl):
        return cs.SpectralLines(sl.name, sl.formula, sl.N, sl.M, pinned(sl.I), pinned(sl.μ), pinned(sl.A), pinned(sl.ν),
                                pinned(sl.S), pinned(sl.γa), pinned(sl.γs), pinned(sl.Epp), pinned(sl.na))

    e2e_gases = []
    for sl, C in my_gases:
        pl = pinned_lines(sl)
        if hasattr(sl, "grid_range"):
            pl.grid_range = sl.grid_range
        e2e_gases.append((pl, C))
    νs_pin, wts_pin = pinned(νs), pinned(wts)
    Fhost = torch.empty(2 * nlev, dtype=torch.float64).pin_memory()
    h2d_step = 0
    for sl, _ in e2e_gases:
        niso, ncheb, cheb, has = sl.cheb_table()
        h2d_step += sum(a.nbytes for a in (sl.ν, sl.S, sl.γa, sl.γs, sl.Epp, sl.na, sl.μ)) + len(sl.I) * 2
        h2d_step += np.asarray(ncheb).nbytes + np.asarray(cheb).nbytes
        h2d_step += nlev * 8 * 8                        # LevelParams block of cs_sigma_add_lines
    h2d_step += 2 * νs_pin.nbytes                       # cs_sigma_create: ν and its trapezoid weights
    h2d_step += wts_pin.nbytes                          # global trapezoid weights of cs_fluxes_device
    h2d_step += (2 * nlev + wl["nlob"] * (nlev - 1) + wl["nlob"] + 3 * wl["nstream"]) * 8   # per-level tables of K6
    d2h_step = Fhost.numel() * 8

    def e2e_step():
        """the call a user makes, from HOST buffers: upload lines + ν, Σ by exact line-by-line gas, fluxes, read F"""
        for sl, _ in e2e_gases:
            sl.__dict__.pop("_dev", None)               # force a fresh cs_lines_upload every step
        gases = [cs.LineGas(sl, C, νs_pin, "voigt", cut, ctx=ctx) for sl, C in e2e_gases]
        A = cs.UnifiedAbsorber(*gases)
        wsx = cs.SigmaWorkspace(νs_pin, nlev, ctx)
        A.sigma_nodes(wsx, Tn, Pn)
        check(lib().cs_fluxes_device(wsx.h, nlev, ptr(Pn), wl["nlob"], ptr(w), ptr(μn), ptr(Tn), wl["g"], None, None,
                                     0.841, wl["nstream"], ptr(m), ptr(W), ptr(wts_pin), dF.data_ptr()))
        if world > 1:
            dist.all_reduce(dF)
        Fhost.copy_(dF, non_blocking=False)
        return Fhost

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        timed_step()
    barrier()
    fp64_peak = ctx.fp64_peak(20000)   # burst DFMA rate of this device [FLOP/s]

    # ---- device-timed region (inputs resident in HBM)
    barrier()
    l0 = ctx.launches()
    tt0 = ctx.timers_total()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            timed_step()
        ev1.record()
        barrier()
        dt = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launches() - l0
    tt1 = ctx.timers_total()
    timers_acc = {k: tt1[k] - tt0[k] for k in tt1}
    tmax = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    F = dF.cpu().numpy()
    olr = float(F[0])

    # ---- end-to-end through the public API with host buffers: the same number of steps, H2D and D2H inside
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    dte = time.perf_counter() - t0
    te = torch.tensor([dte], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dte = float(te.item()) / args.steps
    Fe = Fhost.numpy().copy()

    ms_step = dt / args.steps * 1e3
    value = total_evals / (dt / args.steps)
    ls_ms = timers_acc["linesum"] / args.steps            # per step, summed over this rank's line-sum launches
    n_ls_launch = len(dls)
    achieved = FLOP_PER_EVAL * my_evals / (ls_ms * 1e-3) / 1e12 if ls_ms > 0 else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(wl, world, total_evals),
        "olr_w_m2": olr, "olr_spectra_per_s": 1.0 / (dt / args.steps),
        "e2e": {"value": total_evals / dte, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                "ms_per_step": dte * 1e3, "steps": args.steps, "host_memory": "pinned",
                "max_rel_diff_fluxes_vs_resident": float(np.max(np.abs(Fe - F) / np.maximum(np.abs(F), 1e-300)))},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": {"bound": "fp64", "kernel": "Voigt line sum = line_sum_kernel<VOIGT, COLD> + far_fold_kernel<4, 32> (two launches per gas)",
                     "achieved": achieved, "peak": fp64_peak / 1e12,
                     "unit": "TFLOP/s", "frac": (achieved / (fp64_peak / 1e12)) if achieved else None,
                     "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH if wl["name"] == "c2" and world == 1 else None,
                     "peak_source": "cs_fp64_peak DFMA microbenchmark run in this process (FP64 is not in MEASURED_PEAKS.json; "
                                    "recorded with clocks in profiles/r2_fp64_peak.json)",
                     "flop_per_eval": FLOP_PER_EVAL, "kernel_ms_per_step": ls_ms, "launches_per_step": 2 * n_ls_launch,
                     "kernel_share_of_step": ls_ms / ms_step if ms_step > 0 else None,
                     "rt_kernel_ms_per_step": timers_acc["rt"] / args.steps,
                     "prep_kernel_ms_per_step": timers_acc["prep"] / args.steps},
    }

    # ---- same workload with the far-field expansion switched on (opt-in mode, clearsky_b200.h: CS_FARFIELD_EXPANSION);
    # reported beside the headline, never instead of it
    if not args.no_expansion:
        try:
            ctx.set_farfield("expansion")
            for _ in range(2):
                timed_step()
            barrier()
            tx0 = ctx.timers_total()
            ex0, ex1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ex0.record()
            for _ in range(args.steps):
                timed_step()
            ex1.record()
            barrier()
            tx1 = ctx.timers_total()
            dtx = torch.tensor([ex0.elapsed_time(ex1) * 1e-3], dtype=torch.float64, device=f"cuda:{local}")
            if world > 1:
                dist.all_reduce(dtx, op=dist.ReduceOp.MAX)
            dtx = float(dtx.item()) / args.steps
            Fx = dF.cpu().numpy()
            Σx = ws.read()
            ctx.set_farfield("direct")
            timed_step()
            Σd = ws.read()
            dσ = torch.tensor([float(np.max(np.abs(Σx - Σd) / np.maximum(Σd, 1e-300)))], dtype=torch.float64, device=f"cuda:{local}")
            del Σx, Σd
            if world > 1:
                dist.all_reduce(dσ, op=dist.ReduceOp.MAX)
            nz = np.abs(F) > 0
            line["farfield_expansion"] = {
                "value": total_evals / dtx, "unit": UNIT, "ms_per_step": dtx * 1e3, "olr_spectra_per_s": 1.0 / dtx,
                "olr_w_m2": float(Fx[0]), "max_rel_diff_fluxes_vs_direct": float(np.max(np.abs(Fx[nz] - F[nz]) / np.abs(F[nz]))),
                "max_rel_diff_sigma_vs_direct": float(dσ.item()),
                "linesum_kernel_ms_per_step": (tx1["linesum"] - tx0["linesum"]) / args.steps,
                "note": "far-wing lines >= 4 half tile widths away summed through a 20-term local expansion per tile (32-line clusters "
                        "via 18 moments); truncation < 3e-11 per line"}
        except Exception as e:      # the extra section must never cost the headline line
            ctx.set_farfield("direct")
            line["farfield_expansion"] = {"error": repr(e)}

    # ---- CPU baseline + parity sample (rank 0, N = 1 only): the oracle port on the box's host cores over a bounded ν
    # slice of the SAME workload; its cross-sections and fluxes are kept and the GPU is run on the same slice in both
    # far-field modes and compared (1e-9 on cross-sections, 1e-8 on fluxes: the north-star tolerances)
    parity_fail = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.build()
        nth = host_threads()
        run, evals, sample, (j0, n) = cpu_sample(cs, orc, wl, nth, target_evals=args.cpu_evals if args.cpu_evals > 0 else 2.0e10)
        run()
        t0 = time.perf_counter()
        σo, Fo = run()
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": evals / tc, "unit": UNIT, "cores": nth, "kind": "port", "sample": sample,
                                "seconds": tc}
        νp = np.ascontiguousarray(ν[j0:j0 + n])
        wsp = cs.SigmaWorkspace(νp, nlev, ctx)
        ps = {"n_nu": int(n), "levels": int(nlev), "sample": sample, "tol_sigma": 1e-9, "tol_flux": 1e-8}
        keep = σo > 1e-290
        try:
            for mode in ("direct", "expansion"):
                ctx.set_farfield(mode)
                wsp.zero()
                for dl, C in zip(dls, Cs):
                    check(lib().cs_sigma_add_lines(wsp.h, dl.h, cs._lib.CS_VOIGT, ptr(Tn), ptr(Pn), ptr(C), cut))
                Fg = np.empty(2 * nlev), np.empty(nlev)
                Fup, Fdn, Fnet = np.empty(nlev), np.empty(nlev), np.empty(nlev)
                check(lib().cs_fluxes(wsp.h, nlev, ptr(Pn), wl["nlob"], ptr(w), ptr(μn), ptr(Tn), wl["g"], None, None, 0.841,
                                      wl["nstream"], ptr(m), ptr(W), None, None, None, None, ptr(Fup), ptr(Fdn), ptr(Fnet)))
                Σg = wsp.read()
                ps[f"max_rel_sigma_{mode}"] = float(np.max(np.abs(Σg[keep] - σo[keep]) / σo[keep]))
                ef = max(float(np.max(np.abs(Fup - Fo["Fup"]) / np.abs(Fo["Fup"]))),
                         float(np.max(np.abs(Fdn[1:] - Fo["Fdn"][1:]) / np.abs(Fo["Fdn"][1:]))))
                ps[f"max_rel_flux_{mode}"] = ef
            ps["max_rel_flux"] = max(ps["max_rel_flux_direct"], ps["max_rel_flux_expansion"])
        finally:
            ctx.set_farfield("direct")
        ps["ok"] = bool(ps["max_rel_sigma_direct"] <= 1e-9 and ps["max_rel_sigma_expansion"] <= 1e-9 and ps["max_rel_flux"] <= 1e-8)
        line["parity_sample"] = ps
        if not ps["ok"]:
            parity_fail = ps
    if rank == 0:
        print(json.dumps(line, ensure_ascii=False))
    torch.cuda.synchronize()
    _shutdown(dist, world)
    if parity_fail is not None:
        print(f"bench.py: PARITY FAILURE against the oracle on the sample: {parity_fail}", file=sys.stderr)
        sys.exit(3)


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: radiative-convective loop -- `--workload c5`
C5_METRIC, C5_UNIT = "radiative-convective steps/s", "steps/s"
C5_DT = 600.0


def make_c5(cs, small=False):
    """C2 atmosphere and line lists, 12 x 24 opacity tables on C2's ν grid, 51 cell edges, radmul = 2 -> 101 radiative levels"""
    n, nν = (20_000, 30_000) if small else (250_000, 300_000)
    ν = 0.01 * np.arange(1, nν + 1) * (300_000 / nν)
    Pe = cs.pressuregrid(10.0, 1e5, 51)
    Te = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)(Pe)
    return dict(name="c5small" if small else "c5", n=n, ν=ν, Pe=Pe, Te=Te, g=9.8, μ=0.029, cp=1040.0, cs=1e7, nstream=5, nlob=2,
                radmul=2, domain=((140, 320), 12, (5, 1.1e5), 24),
                gases=[(20261018, 2, (0.06, 0.13), 400e-6), (20261019, 1, (0.10, 0.50), 1e-3)])


def c5_config(wl, world, nrad):
    return {"workload": wl["name"], "lines": 2 * wl["n"], "n_nu": len(wl["ν"]), "cells": len(wl["Pe"]), "radiative_levels": int(nrad),
            "table_nodes": "12x24", "nstream": wl["nstream"], "nlobatto": wl["nlob"], "dt_s": C5_DT,
            "absorber": "AcceleratedAbsorber (frozen, as in the reference's heating!)", "parallelism": f"nu-slices x{world}",
            "l2": "per-step working set (stored transmittances 1.45 GB + Planck scratch 0.24 GB) exceeds the 126 MB L2"}


def c5_cpu_sample(cs, orc, wl, Pr, P, T0, lnσe, i0, n, nthreads):
    """one heating!/step! on the host cores: the oracle's AcceleratedAbsorber evaluation at the radiative levels is static,
    a step is orc.fluxes on the ν sample + the O(np) column arithmetic"""
    ν = np.ascontiguousarray(wl["ν"][i0:i0 + n])
    σr = orc.accel_nodes(np.log(wl["Pe"]), np.ascontiguousarray(lnσe[:, i0:i0 + n]), Pr)
    m, W = cs.streamnodes(wl["nstream"])
    _, w = cs.lobattonodes(wl["nlob"])
    μn = np.full((len(Pr) - 1, wl["nlob"]), wl["μ"])
    Pe = wl["Pe"]

    def step(T, wts=None):
        Tlev = cs.AtmosphericProfile(P, T)(Pr)
        f = orc.fluxes(ν, Pr, wl["nlob"], w, μn, Tlev, σr, wl["g"], None, None, 0.841, wl["nstream"], m, W, nthreads=nthreads, full=False)
        R = -cs.AtmosphericProfile(Pr, f["Fnet"])(Pe)
        H = np.empty(len(Pe))
        H[:-1] = (wl["g"] / wl["cp"]) * (R[:-1] - R[1:]) / (Pe[1:] - Pe[:-1])
        H[-1] = R[-1] / wl["cs"]
        return T + C5_DT * H, H, f

    return step


def run_c5_reference(args):
    import clearsky_b200 as cs
    from oracle import oracle as orc
    if int(os.environ.get("RANK", "0")) != 0:
        return
    orc.build()
    wl = make_c5(cs, args.workload == "c5small")
    nth = host_threads()
    # any smooth positive ln sigma serves the timing: the oracle's cost per step does not depend on the values
    rng = np.random.default_rng(1)
    n = min(len(wl["ν"]), 30_000)
    i0 = (len(wl["ν"]) - n) // 2
    lnσe = np.full((len(wl["Pe"]), len(wl["ν"])), -60.0)
    lnσe[:, i0:i0 + n] += rng.uniform(-3, 3, (1, n))
    rcm = _HostColumn(cs, wl)
    step = c5_cpu_sample(cs, orc, wl, rcm.Pr, rcm.P, rcm.T, lnσe, i0, n, nth)
    T = rcm.T.copy()
    for _ in range(args.warmup):
        T, _, _ = step(T)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        T, _, _ = step(T)
    dt = (time.perf_counter() - t0) / args.steps * (len(wl["ν"]) / n)      # scaled to the full ν grid
    sample = f"ν slice of {n} points of {len(wl['ν'])} x {len(rcm.Pr)} radiative levels, time scaled by the point count"
    cfg = c5_config(wl, args.gpus, len(rcm.Pr))
    cfg["sample"] = sample
    line = {"impl": "reference", "metric": C5_METRIC, "value": 1.0 / dt, "unit": C5_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": 1.0 / dt, "unit": C5_UNIT, "cores": nth, "kind": "port", "sample": sample},
            "e2e": {"value": 1.0 / dt, "unit": C5_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line, ensure_ascii=False))


class _HostColumn:
    """grid bookkeeping of RCM(...) (radiative_convective.jl:42-103) without any device object"""

    def __init__(self, cs, wl):
        Pe, Te, radmul = wl["Pe"], wl["Te"], wl["radmul"]
        n = len(Pe)
        self.P, self.T = np.empty(n), np.empty(n)
        self.P[:-1], self.T[:-1] = (Pe[:-1] + Pe[1:]) / 2, (Te[:-1] + Te[1:]) / 2
        self.P[-1], self.T[-1] = Pe[-1], Te[-1]
        Pr = np.empty(radmul * (n - 1) + 1)
        Pr[: n - 1] = Pe[:-1]
        i = n - 1
        for j in range(2, radmul + 1):
            Pr[i: i + n - 1] = ((j - 1) * Pe[:-1] + (radmul - j + 1) * Pe[1:]) / radmul
            i += n - 1
        Pr[-1] = Pe[-1]
        self.Pr = np.sort(Pr)


def _shutdown(dist, world, grace=20.0):
    """leave torch.distributed without ever hanging the launch: stdout is flushed first, and a watchdog ends the process if
    destroy_process_group does not return (the JSON line is already out)"""
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    t = threading.Timer(grace, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


def run_c5(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import clearsky_b200 as cs
    from clearsky_b200 import sharding
    from clearsky_b200._lib import check, f64, lib, ptr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx = cs.Context(local, stream=stream.cuda_stream)
    wl = make_c5(cs, args.workload == "c5small")
    ν, Pe, Te = wl["ν"], wl["Pe"], wl["Te"]
    edges = balanced_slices(np.ones(len(ν)), world)          # a step costs the same at every wavenumber
    i0, i1 = edges[rank], edges[rank + 1]
    νs = np.ascontiguousarray(ν[i0:i1])
    wts = f64(trapz_weights(ν)[i0:i1])
    Ω = cs.AtmosphericDomain(*wl["domain"])
    t0 = time.perf_counter()
    gases = []
    for seed, M, rng, Cg in wl["gases"]:
        sl = sharding.slice_lines(synthetic_lines(cs, wl["n"], seed, M, rng), νs[0], νs[-1], 25.0, grid=(ν[0], ν[-1]))
        gases.append(cs.Gas(sl, Cg, νs, Ω, ctx=ctx))
    A = cs.AcceleratedAbsorber(Te, Pe, *gases, ctx=ctx)
    ctx.synchronize()
    t_setup = time.perf_counter() - t0
    col = _HostColumn(cs, wl)
    Pr, nrad, npc = f64(col.Pr), len(col.Pr), len(Pe)
    m, W = (f64(x) for x in cs.streamnodes(wl["nstream"]))
    wl_w = f64(cs.lobattonodes(wl["nlob"])[1])
    ws = cs.SigmaWorkspace(νs, nrad, ctx)
    A.sigma_nodes(ws, f64(np.full(nrad, 250.0)), Pr)
    μn = f64(np.full((nrad - 1, wl["nlob"]), wl["μ"]))
    cp = f64(np.full(npc - 1, wl["cp"]))
    h = C.c_void_p()
    check(lib().cs_rcm_create(ws.h, npc, ptr(f64(Pe)), ptr(f64(col.P)), ptr(f64(col.T)), ptr(cp), wl["cs"], nrad, ptr(Pr), wl["nlob"],
                              ptr(wl_w), ptr(μn), wl["g"], None, None, 0.841, wl["nstream"], ptr(m), ptr(W), ptr(wts), C.byref(h)))
    dF = torch.zeros(2 * nrad, dtype=torch.float64, device=dev)
    T0 = f64(col.T)

    # the step's only exchange: 2*nrad doubles.  Default for N > 1: fused into the step's tail kernel through peer-memory
    # mailboxes (CUDA IPC between the ranks, stores over NVLink, cs_rcm_enqueue_step_peer); --c5-collective nccl keeps the
    # NCCL all-reduce between the flux kernels and the column update
    peer = world > 1 and args.c5_collective == "peer"
    opened = []
    if peer:
        box, nb = C.c_void_p(), C.c_int64()
        check(lib().cs_rcm_peer_mailbox(h, world, C.byref(box), C.byref(nb)))
        hb = (C.c_uint8 * 64)()
        check(lib().cs_ipc_export(box, hb))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(hb))
        ptrs = (C.c_void_p * world)()
        for q in range(world):
            if q == rank:
                ptrs[q] = box.value
            else:
                pq = C.c_void_p()
                check(lib().cs_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(handles[q]), C.byref(pq)))
                ptrs[q] = pq.value
                opened.append(pq)
        check(lib().cs_rcm_peer_connect(h, rank, world, ptrs))
        dist.barrier()

    def enqueue_step():
        if peer:
            check(lib().cs_rcm_enqueue_step_peer(h, C5_DT))
            return
        check(lib().cs_rcm_enqueue_fluxes(h, dF.data_ptr()))
        if world > 1:
            dist.all_reduce(dF)
        check(lib().cs_rcm_enqueue_update(h, dF.data_ptr(), C5_DT))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def state():
        T, H = np.empty(npc), np.empty(npc)
        Fup, Fdn = np.empty(nrad), np.empty(nrad)
        check(lib().cs_rcm_state(h, ptr(T), ptr(H), None, ptr(Fup), ptr(Fdn), None))
        return T, H, Fup, Fdn

    # one CUDA graph per step: flux kernels + (NCCL all-reduce) + column update; falls back to plain launches
    graph = None
    for _ in range(max(args.warmup, 3)):
        enqueue_step()
    barrier()
    if not args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                enqueue_step()
            graph = g
        except Exception as e:          # capture is an optimisation, never a requirement
            print(f"bench.py: CUDA-graph capture of the step failed ({e!r}); plain launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    run_step = (lambda: graph.replay()) if graph is not None else enqueue_step
    check(lib().cs_rcm_set_temperature(h, ptr(T0)))
    for _ in range(3):
        run_step()
    check(lib().cs_rcm_set_temperature(h, ptr(T0)))
    barrier()
    l0 = ctx.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            run_step()
        ev1.record()
        barrier()
        dt = ev0.elapsed_time(ev1) * 1e-3
    tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item()) / args.steps
    Tend, Hend, Fup, Fdn = state()

    # kernel-only time of the dominant kernel (rcm_rt_kernel), events on the launching stream around it alone
    ke0, ke1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(5):
        ke0.record()
        check(lib().cs_rcm_enqueue_fluxes(h, dF.data_ptr()))
        ke1.record()
        torch.cuda.synchronize()
        kms.append(ke0.elapsed_time(ke1))
    k_ms = float(np.median(kms))
    L, ns = nrad - 1, wl["nstream"]
    alg_bytes = (2 * (ns + 1) * L + 2 * nrad) * 8.0 * len(νs)       # stored tau + transmittances on both sweeps, Planck scratch w+r
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6458.4))

    # ---- end to end through the host API: temperatures in from pinned host memory, state out, every step
    Tpin = torch.from_numpy(T0.copy()).pin_memory()
    check(lib().cs_rcm_set_temperature(h, ptr(T0)))
    barrier()
    t0 = time.perf_counter()
    Th = T0.copy()
    for _ in range(args.steps):
        Tpin.numpy()[:] = Th
        check(lib().cs_rcm_set_temperature(h, Tpin.numpy().ctypes.data_as(C.POINTER(C.c_double))))
        enqueue_step()
        Th, _, _, _ = state()
    barrier()
    dte = time.perf_counter() - t0
    te = torch.tensor([dte], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dte = float(te.item()) / args.steps

    line = {"metric": C5_METRIC, "value": 1.0 / dt, "unit": C5_UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": c5_config(wl, world, nrad),
            "olr_w_m2": float(Fup[0]), "T_surface_K": float(Tend[-1]), "setup_s": t_setup,
            "step_launch": ("cuda-graph replay" if graph is not None else "plain launches") +
                           (" (flux kernels + one tail kernel: peer-memory exchange over NVLink + column update)" if peer
                            else " (flux kernels + all-reduce + column update)"),
            "collective": ("fused into the step's tail kernel (peer-memory mailboxes, CUDA IPC)" if peer else
                           ("NCCL all-reduce of 2*nrad doubles" if world > 1 else "none (one rank)")),
            "e2e": {"value": 1.0 / dte, "unit": C5_UNIT, "h2d_bytes_per_step": int(npc * 8), "d2h_bytes_per_step": int((2 * npc + 2 * nrad) * 8),
                    "ms_per_step": dte * 1e3, "steps": args.steps, "host_memory": "pinned",
                    "max_rel_diff_T_vs_resident": float(np.max(np.abs(Th - Tend) / Tend))},
            "gpu_launches": int(ctx.launches() - l0) if graph is None else int(3 * args.steps),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": "rcm_rt_kernel<5>", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6458.4 GB/s (MEASURED_PEAKS.json absent)",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / (dt * 1e3), "note": "kernel_ms includes the 2 us spectral reduction launched with it"}}

    parity_fail = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # FULL-size parity of the step against the oracle-driven host twin: the GPU's own ln sigma at the cell edges is the
        # input of both (the tables behind it are checked against the oracle on a slice by tools/config_parity.py c5)
        from oracle import oracle as orc
        orc.build()
        nth = host_threads()
        wse = cs.SigmaWorkspace(νs, npc, ctx)
        A.sigma_nodes(wse, f64(Te), f64(Pe))
        lnσe = np.maximum(np.log(wse.read()), np.log(np.finfo(float).tiny))
        step = c5_cpu_sample(cs, orc, wl, col.Pr, col.P, col.T, lnσe, 0, len(ν), nth)
        check(lib().cs_rcm_set_temperature(h, ptr(T0)))
        T = T0.copy()
        eT = eH = eF = 0.0
        tc = []
        for _ in range(3):
            t0 = time.perf_counter()
            T, H, f = step(T)
            tc.append(time.perf_counter() - t0)
            enqueue_step()
            Tg, Hg, Fu, Fd = state()
            eT = max(eT, float(np.max(np.abs(Tg - T) / T)))
            eH = max(eH, float(np.max(np.abs(Hg - H)) / np.max(np.abs(H))))
            eF = max(eF, float(np.max(np.abs(Fu - f["Fup"]) / f["Fup"])), float(np.max(np.abs(Fd[1:] - f["Fdn"][1:]) / f["Fdn"][1:])))
        ps = {"n_nu": len(ν), "levels": nrad, "steps": 3, "sample": "full workload (flux solve + column update; Sigma taken from the GPU)",
              "max_rel_T": eT, "max_rel_heating": eH, "max_rel_flux": eF, "tol_flux": 1e-8, "ok": bool(max(eT, eH, eF) <= 1e-8)}
        line["parity_sample"] = ps
        line["cpu_baseline"] = {"value": 1.0 / min(tc), "unit": C5_UNIT, "cores": nth, "kind": "port",
                                "sample": "full workload: orc.fluxes over all ν at the 101 radiative levels + column update", "seconds": min(tc)}
        if not ps["ok"]:
            parity_fail = ps
    if rank == 0:
        print(json.dumps(line, ensure_ascii=False))
    # the captured step holds the NCCL all-reduce: release the graph before the communicator goes away (destroying the process
    # group under a live graph hung the torchrun launch of this workload), and never let the shutdown outlive the result
    run_step = None
    graph = None
    gc.collect()
    torch.cuda.synchronize()
    if peer:
        late = C.c_int32(0)
        check(lib().cs_rcm_peer_status(h, None, C.byref(late)))
        if late.value:
            print("bench.py: a rank's partial fluxes did not arrive within the time limit of the fused step", file=sys.stderr)
            parity_fail = parity_fail or {"peer_exchange": "timed out"}
        dist.barrier()                   # nobody unmaps a mailbox another rank may still be writing to
        for pq in opened:
            lib().cs_ipc_close(ctx.h, pq)
        dist.barrier()
    lib().cs_rcm_free(h)
    _shutdown(dist, world)
    if parity_fail is not None:
        print(f"bench.py: PARITY FAILURE against the oracle: {parity_fail}", file=sys.stderr)
        sys.exit(3)


def run_single_process(args):
    """configs[1] from ONE host process over N GPUs: the host model the Julia wrapper uses.  nu slices
    balanced by the same cost model, lines uploaded per slice, one thread per device, the library's ncclAllReduce (cs_group.cu)
    on the 2*np integrated fluxes.  A step is wall-clock timed around sh.fluxes() (the call returns after the reduced fluxes are
    on the host), so it is an end-to-end number with resident line lists; the kernel times come from the per-context timers."""
    import clearsky_b200 as cs
    n = max(1, args.gpus)
    assert cs.device_count() >= n, f"--gpus {n} but only {cs.device_count()} device(s) visible"
    wl = make_workload(cs, args.workload)
    ν, P, T = wl["ν"], wl["P"], wl["T"]
    # NCCL prints a version banner on stdout when the communicator comes up (first collective); stdout must carry exactly
    # one JSON line, so file descriptor 1 points at stderr until the warm-up steps are done
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        grp = cs.DeviceGroup(list(range(n)))
        sh = cs.ShardedLineByLine(grp, [(sl, C, "voigt", wl["cut"]) for sl, C in wl["gases"]], ν)
        prof = cs.AtmosphericProfile(P, T)
        total_evals = 0
        for part in sh.parts:
            if part is not None:
                total_evals += sum(g.evals_per_node() for g in part["gases"]) * len(P)
        for _ in range(max(args.warmup, 3)):
            Fup, Fdn, Fnet = sh.fluxes(P, wl["g"], prof, wl["μ"])
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    t0 = [c.timers_total() for c in grp.ctx]
    with ClockSampler(0) as clk:
        w0 = time.perf_counter()
        for _ in range(args.steps):
            Fup, Fdn, Fnet = sh.fluxes(P, wl["g"], prof, wl["μ"])
        dt = (time.perf_counter() - w0) / args.steps
    t1 = [c.timers_total() for c in grp.ctx]
    ls = [(b["linesum"] - a["linesum"]) / args.steps for a, b in zip(t0, t1)]
    cfg = workload_config(wl, n, total_evals)
    cfg["parallelism"] = f"nu-slices x{n}, single process (cs_group: thread per device + library ncclAllReduce)"
    line = {"metric": METRIC, "value": total_evals / dt, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "olr_w_m2": float(Fup[0]), "host_model": "single-process",
            "timing": "host wall clock around the synchronous group call (includes the all-reduce and the D2H of the fluxes)",
            "linesum_kernel_ms_per_device": ls, "slice_imbalance_max_over_mean": max(ls) / (sum(ls) / len(ls)),
            "clocks": clk.summary()}
    grp.close()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c2small", "c5", "c5small"])
    ap.add_argument("--cpu-evals", type=float, default=0.0,
                    help="size of the bounded CPU sample [evals]; 0 = 2e10 for the cpu_baseline leg, and for --impl reference "
                         "a size calibrated so that warm-up + steps take about 150 s")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-expansion", action="store_true", help="skip the extra far-field-expansion measurement")
    ap.add_argument("--no-graph", action="store_true", help="c5: plain launches instead of a CUDA graph per step")
    ap.add_argument("--c5-collective", default="peer", choices=["peer", "nccl"],
                    help="c5, N > 1: exchange of the 2*nrad partial fluxes -- fused into the step's tail kernel over peer memory "
                         "(default) or an NCCL all-reduce between the flux kernels and the column update")
    ap.add_argument("--single-process", action="store_true",
                    help="c2: drive all --gpus N devices from THIS process through the library's own device group (cs_group_*: one "
                         "host thread per device, ncclAllReduce inside the library) instead of one torchrun rank per GPU")
    args = ap.parse_args()
    c5 = args.workload.startswith("c5")
    if args.impl == "reference":
        (run_c5_reference if c5 else run_reference)(args)
    elif args.single_process and not c5:
        run_single_process(args)
    else:
        (run_c5 if c5 else run_ours)(args)


if __name__ == "__main__":
    main()
