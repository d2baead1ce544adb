#!/usr/bin/env python
"""Build compile-time variants of the line-sum kernel (cs_lines.cu only) and time them on C2.

    python tools/k2_variants.py build  NAME="-DCS_LS_FOLD=8 ..." [NAME=...]     # here (no GPU): lib/variants/NAME/*.so
    python tools/k2_variants.py time [NLEV]                                      # on the GPU box: every built variant

Each variant is timed in its own process (CLEARSKY_B200_LIB), direct and expansion mode, line-sum kernel time from the
library's event timers; Sigma is compared with the first variant's (max relative difference) as a sanity check."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "clearsky.jl_b200")
VAR = os.path.join(PKG, "lib", "variants")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
         "--fmad=true"]


def build(specs):
    objs = [os.path.join(PKG, "lib", f) for f in ("cs_api.o", "cs_table.o", "cs_rt.o", "cs_group.o", "cs_par.o")]
    for spec in specs:
        name, _, flags = spec.partition("=")
        d = os.path.join(VAR, name)
        os.makedirs(d, exist_ok=True)
        obj = os.path.join(d, "cs_lines.o")
        subprocess.check_call([NVCC] + FLAGS + flags.split() + ["-c", os.path.join(PKG, "csrc", "cs_lines.cu"), "-o", obj])
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC",
                               "-o", os.path.join(d, "libclearsky_b200.so"), obj] + objs + ["-ldl"])
        os.remove(obj)
        print("built", name, flags)


def worker(nlev):
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import clearsky_b200 as cs
    from clearsky_b200._lib import check, f64, lib, ptr
    wl = bench.make_workload(cs, "c2")
    ctx = cs.default_context()
    ν, P, T = wl["ν"], wl["P"], wl["T"]
    if nlev < len(P):
        k = np.linspace(0, len(P) - 1, nlev).astype(int)
        P, T = P[k], T[k]
    dls = [cs.DeviceLines(sl, ctx) for sl, _ in wl["gases"]]
    ws = cs.SigmaWorkspace(ν, len(P), ctx)
    Cs = [f64(np.full(len(P), C)) for _, C in wl["gases"]]
    out = {}
    for mode in ("direct", "expansion"):
        ctx.set_farfield(mode)
        best = 1e30
        for it in range(4):
            ws.zero()
            t0 = ctx.timers_total()["linesum"]
            for dl, C in zip(dls, Cs):
                check(lib().cs_sigma_add_lines(ws.h, dl.h, 2, ptr(f64(T)), ptr(f64(P)), ptr(C), 25.0))
            ctx.synchronize()
            best = min(best, ctx.timers_total()["linesum"] - t0)
        out[mode] = best
        sub = ws.read()[:, 140000:160000]
        np.save(f"/tmp/k2var_{os.environ['K2_NAME']}_{mode}.npy", sub)
    print(f"{os.environ['K2_NAME']:28s} direct {out['direct']:8.3f} ms   expansion {out['expansion']:8.3f} ms", flush=True)


def time_all(nlev):
    import numpy as np
    names = sorted(os.listdir(VAR))
    # K2_ENVS="CS_LINESUM_NCOLD=0;CS_LINESUM_NCOLD=4": every built variant is also timed under each of these run-time settings
    envs = [e for e in os.environ.get("K2_ENVS", "").split(";") if e]
    runs = [(n, n, {}) for n in names] if not envs else [(n, f"{n}[{e}]", dict([e.split("=", 1)])) for n in names for e in envs]
    for n, label, extra in runs:
        env = dict(os.environ, CLEARSKY_B200_LIB=os.path.join(VAR, n, "libclearsky_b200.so"), K2_NAME=label, **extra)
        subprocess.call([sys.executable, os.path.abspath(__file__), "worker", str(nlev)], env=env)
    names = [label for _, label, _ in runs]
    for mode in ("direct", "expansion"):
        ref = np.load(f"/tmp/k2var_{names[0]}_{mode}.npy")
        for n in names[1:]:
            x = np.load(f"/tmp/k2var_{n}_{mode}.npy")
            print(f"{mode:9s} {n:28s} max rel diff vs {names[0]}: {float(np.max(np.abs(x - ref) / ref)):.2e}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "worker":
        worker(int(sys.argv[2]))
    else:
        time_all(int(sys.argv[2]) if len(sys.argv) > 2 else 101)
