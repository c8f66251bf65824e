"""Sharding of the (spectral bin x column) batch over the GPUs of one box.

Every (spectral point, column) solve is independent, so ranks take contiguous equal slices and the
solve itself needs no communication.  The only exchange is the band-integrated flux
bflx(ncol, nlev, 2) = sum_n w_n F_n (examples/amars_lw.cpp:84-88):

  * sharded by COLUMN (ncol >= world): every rank holds all spectral points of its columns, the
    weighted sum is rank-local, and an all-gather along the column axis assembles the result
    (no reduction crosses NVLink);
  * sharded by SPECTRAL POINT (few columns, e.g. line-by-line): each rank reduces its points locally
    (b200disort_reduce_band) and one all-reduce (sum, FP64) of the small (ncol, nlev, 2) tensor
    finishes the integration -- NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    axis: str     # "col" or "wave"
    start: int
    stop: int
    world: int
    rank: int

    @property
    def size(self):
        return self.stop - self.start


def _split(n, world, rank):
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def partition(nwave, ncol, world, rank):
    """Contiguous slice owned by `rank`: by column when there are enough columns, else by spectral point."""
    if not 0 <= rank < world:
        raise ValueError("rank outside world")
    if ncol >= world:
        return Shard("col", *_split(ncol, world, rank), world, rank)
    if nwave < world:
        raise ValueError(f"cannot shard nwave={nwave}, ncol={ncol} over {world} ranks")
    return Shard("wave", *_split(nwave, world, rank), world, rank)


def take(shard, prop, bc=None, temf=None, weights=None, wave_lower=None, wave_upper=None):
    """Slice the inputs of Disort.forward for this shard (works on numpy arrays and torch tensors)."""
    sl = slice(shard.start, shard.stop)
    if shard.axis == "col":
        return (prop[:, sl], {k: v[:, sl] for k, v in (bc or {}).items()},
                None if temf is None else temf[sl], weights, wave_lower, wave_upper)
    cut = lambda a: None if a is None else a[sl]
    return (prop[sl], {k: v[sl] for k, v in (bc or {}).items()}, temf, cut(weights), cut(wave_lower),
            cut(wave_upper))


def band_flux(shard, bflx_local, ncol_total, group=None):
    """Finish the band integration across ranks.  bflx_local is this rank's weighted sum over its
    spectral points, shape (ncol_local, nlev, 2).  Returns (ncol_total, nlev, 2) on every rank."""
    if shard.world == 1:
        return bflx_local
    if shard.axis == "wave":
        out = bflx_local.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out
    # all-gather along the column axis; slices differ by at most one column, so pad to the largest
    sizes = [b - a for a, b in (_split(ncol_total, shard.world, r) for r in range(shard.world))]
    big = max(sizes)
    pad = torch.zeros((big,) + tuple(bflx_local.shape[1:]), dtype=bflx_local.dtype, device=bflx_local.device)
    pad[:bflx_local.shape[0]] = bflx_local
    out = torch.empty((shard.world * big,) + tuple(bflx_local.shape[1:]), dtype=bflx_local.dtype,
                      device=bflx_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((shard.world, big) + tuple(bflx_local.shape[1:]))
    return torch.cat([out[r, :n] for r, n in enumerate(sizes)], dim=0)
