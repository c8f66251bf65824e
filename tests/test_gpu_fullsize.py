"""BASELINE.json's full-size configurations on one B200: the GCM batch (65 536 columns x 32 g-points x 60 layers,
nstr = 16, 2.1 M solves, 18 GB of optical properties resident in HBM) and the line-by-line case (100 000
wavenumbers x 1 column, nstr = 4).  Checked through size-independent properties (linearity in the beam,
tile determinism, status words) and against the CPU oracle on a random sample of solves at 1e-9."""
import numpy as np
import pytest
import torch

import oracle_lib as O
from pyharp_b200 import Disort, workloads as W

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _sample_check(wl, flx, idx):
    n, j = np.unravel_index(idx, wl.shape[:2])
    prop = wl.prop[n, j][:, None]
    bc = {k: v[n, j][:, None] for k, v in wl.bc.items()}
    ref, _, st = O.forward(prop, bc, None, nstr=wl.nstr, flags=wl.flags, nthreads=8)
    assert not st.any()
    got = flx[n, j][:, None]
    err = np.max(np.abs(got - ref) / np.max(np.abs(ref), axis=(2, 3), keepdims=True))
    assert err < RTOL, err


def test_gcm_full_size(cuda_device):
    ncol_tile, ntile, nwave, nlyr = 4096, 16, 32, 60
    wl = W.make("gcm", nwave=nwave, ncol=ncol_tile, nlyr=nlyr)
    ncol = ncol_tile * ntile
    prop = torch.from_numpy(wl.prop).to(cuda_device).repeat(1, ntile, 1, 1)      # (32, 65536, 60, 18): 18.1 GB
    bc = {k: torch.from_numpy(v).to(cuda_device).repeat(1, ntile) for k, v in wl.bc.items()}
    assert prop.shape == (nwave, ncol, nlyr, 18) and prop.numel() * 8 > 18e9
    op = wl.options().ncol(ncol)
    ds = Disort(op)
    flx = ds.forward(prop, bc)
    assert flx.shape == (nwave, ncol, nlyr + 1, 2)
    assert int(ds.status().abs().sum()) == 0 and bool(torch.isfinite(flx).all())
    # every tile solved the same atmospheres: bit-identical results wherever they ran on the GPU
    tiles = flx.view(nwave, ntile, ncol_tile, nlyr + 1, 2)
    assert bool((tiles == tiles[:, :1]).all())
    # beam-only problem: fluxes are linear in the beam.  Not to the last bit: among 2.1 M solves x 60 layers x 8
    # eigenvalues some 1/mu0 fall within ~1e-6 of an eigenvalue k, where the beam particular solution ~ 1/(k^2 - 1/mu0^2)
    # cancels against the homogeneous one and rounding is amplified to ~1e-10 (DISORT has the same conditioning; it only
    # warns in c_upbeam).  The criterion is the parity tolerance, relative to each solve's largest flux.
    bc2 = dict(bc)
    bc2["fbeam"] = bc["fbeam"] * 3.0
    flx3 = ds.forward(prop, bc2)
    scale = flx.abs().amax(dim=(2, 3), keepdim=True)
    assert float(((flx3 / 3.0 - flx).abs() / scale).max()) < RTOL
    # the direct beam at the top of the atmosphere is mu0 * fbeam
    torch.testing.assert_close(flx[:, :, nlyr, 1], bc["umu0"] * bc["fbeam"], rtol=1e-14, atol=0)
    rng = np.random.default_rng(1)
    _sample_check(wl, flx[:, :ncol_tile].cpu().numpy(), rng.choice(nwave * ncol_tile, 48, replace=False))


def test_lbl_full_size(cuda_device):
    wl = W.make("lbl")  # 100 000 x 1 x 40, nstr = 4
    assert wl.shape[:3] == (100000, 1, 40)
    ds = Disort(wl.options())
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()})
    assert int(ds.status().abs().sum()) == 0
    f = flx.cpu().numpy()
    rng = np.random.default_rng(2)
    _sample_check(wl, f, rng.choice(100000, 400, replace=False))
    # band integration on the device equals the host sum (examples/amars_lw.cpp:88)
    w = torch.full((100000,), 1e-5, dtype=torch.float64)
    got = ds.reduce_band(flx, w).cpu()
    want = (flx.cpu() * w.view(-1, 1, 1, 1)).sum(0)
    torch.testing.assert_close(got, want, rtol=1e-12, atol=0)


def test_lw_large_batch_host_path(cuda_device):
    """131 072 thermal solves through the pipelined host-pointer path (pageable AND pinned buffers) equal the
    device-pointer path bit for bit."""
    wl = W.make("amars_lw", nwave=32, ncol=4096, nlyr=80)
    ds = Disort(wl.options())
    t = lambda a, d: torch.from_numpy(a).to(d)
    dev = ds.forward(t(wl.prop, cuda_device), {k: t(v, cuda_device) for k, v in wl.bc.items()}, t(wl.temf, cuda_device)).cpu()
    pageable = ds.forward(t(wl.prop, "cpu"), {k: t(v, "cpu") for k, v in wl.bc.items()}, t(wl.temf, "cpu"))
    pinned = ds.forward(t(wl.prop, "cpu").pin_memory(), {k: t(v, "cpu").pin_memory() for k, v in wl.bc.items()},
                        t(wl.temf, "cpu").pin_memory())
    assert torch.equal(dev, pageable) and torch.equal(dev, pinned)
    assert bool(torch.isfinite(dev).all())
