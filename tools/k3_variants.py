#!/usr/bin/env python
"""Build compile-time variants of the fused table fit (cs_table.cu only) and time them on the C4 fit (50 x 50 nodes, 1e6 nu).

    python tools/k3_variants.py build NAME="-DCS_FF_THREADS=416" [NAME=@path/to/other_cs_table.cu ...]   # here (no GPU)
    python tools/k3_variants.py time [NNU]                                                                # on the GPU box

A variant given as NAME=@file compiles that file in place of csrc/cs_table.cu (e.g. `git show HEAD:...cs_table.cu > /tmp/x.cu`).
Each variant is timed in its own process (CLEARSKY_B200_LIB); the coefficients of a slice are compared with the first variant's."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "clearsky.jl_b200")
VAR = os.path.join(PKG, "lib", "variants_k3")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
         "--fmad=true", "-I", os.path.join(PKG, "csrc")]


def build(specs):
    objs = [os.path.join(PKG, "lib", f) for f in ("cs_api.o", "cs_lines.o", "cs_rt.o", "cs_group.o", "cs_par.o")]
    for spec in specs:
        name, _, flags = spec.partition("=")
        src = os.path.join(PKG, "csrc", "cs_table.cu")
        if flags.startswith("@"):
            src, flags = flags[1:], ""
        d = os.path.join(VAR, name)
        os.makedirs(d, exist_ok=True)
        obj = os.path.join(d, "cs_table.o")
        subprocess.check_call([NVCC] + FLAGS + flags.split() + ["-c", src, "-o", obj])
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC",
                               "-o", os.path.join(d, "libclearsky_b200.so"), obj] + objs + ["-ldl"])
        os.remove(obj)
        print("built", name, flags or src)


def worker(nν):
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import clearsky_b200 as cs
    ctx = cs.default_context()
    ctx.set_farfield("expansion")                       # keeps the bake's line sum short; the fit does not depend on it
    sl = bench.synthetic_lines(cs, 250_000, 20261018, 2, (0.06, 0.13))
    ν = (3000.0 / nν) * np.arange(1, nν + 1)
    Ω = cs.AtmosphericDomain((150, 320), 50, (5, 1.1e5), 50)
    best = 1e30
    for it in range(3):
        gas = cs.Gas(sl, 400e-6, ν, Ω)
        best = min(best, gas.timers["table_fit"])
        if it < 2:
            del gas
    T = np.linspace(160.0, 310.0, 7)
    P = np.geomspace(10.0, 1e5, 7)
    np.save(f"/tmp/k3var_{os.environ['K3_NAME']}.npy", np.asarray(gas.rawσ(T, P))[:, ::997])
    print(f"{os.environ['K3_NAME']:24s} fit {best:8.3f} ms", flush=True)


def time_all(nν):
    import numpy as np
    names = sorted(os.listdir(VAR))
    for n in names:
        env = dict(os.environ, CLEARSKY_B200_LIB=os.path.join(VAR, n, "libclearsky_b200.so"), K3_NAME=n)
        subprocess.call([sys.executable, os.path.abspath(__file__), "worker", str(nν)], env=env)
    try:
        ref = np.load(f"/tmp/k3var_{names[0]}.npy")
        for n in names[1:]:
            x = np.load(f"/tmp/k3var_{n}.npy")
            print(f"{n:24s} max rel diff of table values vs {names[0]}: {float(np.max(np.abs(x - ref) / np.abs(ref))):.2e}")
    except FileNotFoundError:
        pass


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "worker":
        worker(int(sys.argv[2]))
    else:
        time_all(int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000)
