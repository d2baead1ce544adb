"""world_size-2 gloo test of the N>1 host path (no GPU): ν slices balanced by evaluation count, per-rank partial
spectral integrals with the global trapezoid weights, one all-reduce of the 2*np fluxes (the only collective on
the path).  The per-rank monochromatic fluxes come from the CPU oracle here; on the GPU box the same slicing feeds
cs_fluxes_device + NCCL (bench.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import DATA, ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import bench
    import clearsky_b200 as cs
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    co2 = cs.SpectralLines.from_file(os.path.join(DATA, "CO2.par.gz"), νmin=500, νmax=900)
    ν = np.linspace(550.0, 850.0, 601)
    P = cs.pressuregrid(10.0, 1e5, 9)
    Γ = cs.DryAdiabat(288.0, 1e5, 1040.0, 0.029, Ptropo=1e4)
    T = Γ(P)
    cnt = bench.per_point_counts(ν, co2.ν, 25.0)
    e = bench.balanced_slices(cnt, world)
    a, b = e[rank], e[rank + 1]
    w = bench.trapz_weights(ν)[a:b]
    m, W = cs.streamnodes(5)
    x, wl = cs.lobattonodes(2)
    μn = np.full((len(P) - 1, 2), 0.029)
    σ = 400e-6 * orc.xsec(orc.VOIGT, co2, ν[a:b], T, P, 400e-6 * P, 25.0)
    f = orc.fluxes(ν[a:b], P, 2, wl, μn, T, σ, 9.8, None, None, 0.841, 5, m, W)
    part = np.concatenate([w @ f["Mup"], w @ f["Mdn"]])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    if rank == 0:
        σf = 400e-6 * orc.xsec(orc.VOIGT, co2, ν, T, P, 400e-6 * P, 25.0)
        ref = orc.fluxes(ν, P, 2, wl, μn, T, σf, 9.8, None, None, 0.841, 5, m, W)
        np.save(out, np.stack([t.numpy(), np.concatenate([ref["Fup"], ref["Fdn"]])]))
    dist.destroy_process_group()


def test_two_rank_nu_sharding(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, ref = np.load(out)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-12
