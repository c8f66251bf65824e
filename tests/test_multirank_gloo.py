"""world_size-2 test of the multi-GPU host logic on CPU (gloo): sharding by column and by spectral
point, local band integration, and the one collective that finishes it.  The per-rank solve is done
by the CPU oracle here (the CUDA library needs a GPU); what is under test is pyharp_b200.shard."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from pyharp_b200 import shard as S, workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solve(prop, bc, temf, wl, wlo, whi):
    flx, _, st = O.forward(np.ascontiguousarray(prop), {k: np.ascontiguousarray(v) for k, v in bc.items()},
                           None if temf is None else np.ascontiguousarray(temf), nstr=wl.nstr, flags=wl.flags,
                           wave_lower=wlo, wave_upper=whi)
    assert not st.any()
    return flx


def _worker(rank, world, port, name, kw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = W.make(name, **kw)
    nwave, ncol = wl.shape[:2]
    weights = np.linspace(0.5, 1.5, nwave) / nwave
    sh = S.partition(nwave, ncol, world, rank)
    prop, bc, temf, w, wlo, whi = S.take(sh, wl.prop, wl.bc, wl.temf, weights, wl.wave_lower, wl.wave_upper)
    flx = _solve(prop, bc, temf, wl, wlo, whi)
    local = torch.from_numpy((flx * w[:, None, None, None]).sum(0))
    out = S.band_flux(sh, local, ncol)
    if rank == 0:
        q.put((sh.axis, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,kw,axis", [("amars_lw", dict(nwave=6, ncol=5, nlyr=12), "col"),
                                          ("lbl", dict(nwave=9, ncol=1, nlyr=10), "wave")])
def test_two_rank_band_flux(name, kw, axis):
    wl = W.make(name, **kw)
    nwave = wl.shape[0]
    weights = np.linspace(0.5, 1.5, nwave) / nwave
    full = _solve(wl.prop, wl.bc, wl.temf, wl, wl.wave_lower, wl.wave_upper)
    want = (full * weights[:, None, None, None]).sum(0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, kw, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got_axis, got = q.get(timeout=10)
    assert got_axis == axis
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-300)


def test_partition_covers_everything():
    for nwave, ncol, world in ((32, 65536, 8), (100000, 1, 8), (500, 1, 4), (32, 7, 4), (10, 3, 2)):
        seen = []
        for r in range(world):
            sh = S.partition(nwave, ncol, world, r)
            seen.append((sh.axis, sh.start, sh.stop))
        assert len({a for a, _, _ in seen}) == 1
        n = ncol if seen[0][0] == "col" else nwave
        assert seen[0][1] == 0 and seen[-1][2] == n
        assert all(seen[i][2] == seen[i + 1][1] for i in range(world - 1))
        sizes = [b - a for _, a, b in seen]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        S.partition(2, 1, 4, 0)
