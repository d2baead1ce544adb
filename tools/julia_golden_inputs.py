#!/usr/bin/env python
"""Write the synthetic-line inputs of tools/julia_golden.jl: the part of BASELINE configs[1]'s two line lists that a
1500-point slice of its wavenumber grid can see, as real 160-column HITRAN records (writepar), so the UNMODIFIED
reference's `readpar` reads exactly the lines this repo's generator makes.

    python tools/julia_golden_inputs.py          # -> tests/golden/ref_inputs/c2slice_{CO2,H2O}.par.gz + c2slice.json

Host-side code only (no GPU call, no oracle).  The outputs are committed; rerun only if the generator changes.
"""
import gzip
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)

POINTS, WHERE, CUT, MARGIN = 1500, 0.5, 25.0, 0.5


def take(cs, sl, lo, hi):
    k = (sl.ν >= lo) & (sl.ν <= hi)
    f = lambda a: np.ascontiguousarray(a[k])
    return cs.SpectralLines(sl.name, sl.formula, int(k.sum()), sl.M, f(sl.I), f(sl.μ), f(sl.A), f(sl.ν), f(sl.S), f(sl.γa),
                            f(sl.γs), f(sl.Epp), f(sl.na))


def main():
    import bench
    import clearsky_b200 as cs
    out = os.path.join(ROOT, "tests", "golden", "ref_inputs")
    os.makedirs(out, exist_ok=True)
    wl = bench.make_workload(cs, "c2")
    ν = wl["ν"]
    i0 = int((len(ν) - POINTS) * WHERE)
    νs = ν[i0:i0 + POINTS]
    meta = {"first_index": i0, "n_nu": POINTS, "nu_first": float(νs[0]), "nu_step": 0.01, "cut": CUT, "gases": {}}
    for (sl, C), name in zip(wl["gases"], ("CO2", "H2O")):
        sub = take(cs, sl, νs[0] - CUT - MARGIN, νs[-1] + CUT + MARGIN)
        tmp = os.path.join(out, f"c2slice_{name}.par")
        cs.writepar(tmp, sub)
        with open(tmp, "rb") as f, gzip.GzipFile(tmp + ".gz", "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
        os.remove(tmp)
        meta["gases"][name] = {"lines": sub.N, "C": C}
        print(name, sub.N, "lines ->", tmp + ".gz")
    with open(os.path.join(out, "c2slice.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
