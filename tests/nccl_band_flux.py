"""Spectral-point and column sharding over real GPUs with NCCL: every rank solves its slice with the
CUDA library, integrates its band flux on the device (b200disort_reduce_band) and the one collective of the path
(all-reduce for spectral-point shards, all-gather for column shards; pyharp_b200.shard.band_flux) finishes it.
Rank 0 compares with the whole problem solved on its own GPU.  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_band_flux.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyharp_b200 import Disort, shard as S, workloads as W  # noqa: E402


def solve(wl, prop, bc, temf, wlo, whi, dev):
    nw, nc = prop.shape[:2]
    op = wl.options().nwave(nw).ncol(nc)
    if wlo is not None:
        op.wave_lower(list(wlo)).wave_upper(list(whi))
    ds = Disort(op)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    flx = ds.forward(t(prop), {k: t(v) for k, v in bc.items()}, t(temf))
    assert int(ds.status().abs().sum()) == 0
    return ds, flx


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    worst = 0.0
    for name, kw in (("lbl", dict(nwave=4001, ncol=1, nlyr=40)), ("amars_lw", dict(nwave=16, ncol=301, nlyr=40)),
                     ("amars_sw", dict(nwave=500, ncol=1, nlyr=40))):
        wl = W.make(name, **kw)
        nwave, ncol = wl.shape[:2]
        weights = np.linspace(0.5, 1.5, nwave) / nwave
        sh = S.partition(nwave, ncol, world, rank)
        prop, bc, temf, w, wlo, whi = S.take(sh, wl.prop, wl.bc, wl.temf, weights, wl.wave_lower, wl.wave_upper)
        ds, flx = solve(wl, prop, bc, temf, wlo, whi, dev)
        local = ds.reduce_band(flx, torch.from_numpy(np.ascontiguousarray(w)).to(dev))
        out = S.band_flux(sh, local, ncol)
        if rank == 0:
            ds0, full = solve(wl, wl.prop, wl.bc, wl.temf, wl.wave_lower, wl.wave_upper, dev)
            want = ds0.reduce_band(full, torch.from_numpy(weights).to(dev))
            err = float((out - want).abs().max() / want.abs().max())
            worst = max(worst, err)
            print(f"{name}: sharded by {sh.axis} over {world} GPUs, band flux {tuple(out.shape)}, max rel diff {err:.2e}", flush=True)
            assert err < 1e-12, (name, err)
        dist.barrier()
    if rank == 0:
        print(f"NCCL_BAND_FLUX_OK worst {worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
