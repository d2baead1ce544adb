#!/usr/bin/env python
"""Build libclearsky_b200.so (sm_100a only) in-tree with nvcc.

    python clearsky.jl_b200/build.py [--force] [--verbose]

The shared object lands in clearsky.jl_b200/lib/ (git-ignored, travels to the GPU box with gpurun).
There is deliberately no other architecture and no CPU fallback.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OUT = os.path.join(LIB, "libclearsky_b200.so")
SOURCES = ["cs_api.cu", "cs_lines.cu", "cs_table.cu", "cs_rt.cu", "cs_group.cu", "cs_par.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--fmad=true"] + os.environ.get("CS_NVCC_EXTRA", "").split()


def _deps():
    hdrs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "clearsky_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(SRC, s)
        obj = os.path.join(LIB, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + _deps()):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr
    with ThreadPoolExecutor(max_workers=4) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xcompiler", "-fPIC", "-o", OUT] + objs + ["-ldl"]
        run(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
