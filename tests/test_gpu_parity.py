"""Parity of the CUDA path (through the C ABI) against the CPU oracle: relative tolerance 1e-9
on fluxes (BASELINE.json north_star), measured against the largest flux of the solve."""
import numpy as np
import pytest
import torch

import oracle_lib as O
from pyharp_b200 import Disort, workloads as W

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _oracle(wl):
    ref, _, st = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags,
                           wave_lower=wl.wave_lower, wave_upper=wl.wave_upper, nthreads=8)
    assert not st.any()
    return ref


def _relerr(a, ref):
    scale = np.max(np.abs(ref), axis=(2, 3), keepdims=True)
    return float(np.max(np.abs(a - ref) / scale))


CASES = [
    ("test1a", dict()),
    ("amars_sw", dict(nwave=64, ncol=3, nlyr=40)),
    ("amars_lw", dict(nwave=32, ncol=2, nlyr=40)),
    ("amars_lw", dict(nwave=32, ncol=2, nlyr=80)),
    ("amars_lw", dict(nwave=16, ncol=2, nlyr=80, variant=1)),
    ("lbl", dict(nwave=2000, ncol=1, nlyr=40)),
    ("gcm", dict(nwave=4, ncol=64, nlyr=60)),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(CASES)])
@pytest.mark.parametrize("where", ["host", "device"])
def test_flux_parity(cuda_device, name, kw, where):
    wl = W.make(name, **kw)
    ref = _oracle(wl)
    ds = Disort(wl.options())
    dev = "cpu" if where == "host" else cuda_device
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf))
    assert flx.device.type == ("cpu" if where == "host" else "cuda")
    assert flx.shape == ref.shape
    err = _relerr(flx.cpu().numpy(), ref)
    assert err < RTOL, err
    assert int(ds.status().abs().sum()) == 0
    assert ds.stats()["kernel_launches"] >= 1


def test_golden_test1a(cuda_device):
    """DISORT test problem 1a as tests/test_disort.cpp sets it up."""
    wl = W.make("test1a")
    ds = Disort(wl.options())
    flx = ds.forward(torch.from_numpy(wl.prop), {k: torch.from_numpy(v) for k, v in wl.bc.items()}).numpy()
    for n in range(flx.shape[0]):
        assert abs(flx[n, 0, 0, 0]) < 1e-12                       # flup at the surface
        assert abs(flx[n, 0, 0, 1] / 2.3778490806 - 1) < 1e-9     # rfldir + rfldn at the surface
        assert abs(flx[n, 0, 1, 0] / 7.9945097492e-02 - 1) < 1e-9 # flup at the top
        assert abs(flx[n, 0, 1, 1] / np.pi - 1) < 1e-12


def test_bad_input_raises(cuda_device):
    wl = W.make("amars_sw", nwave=4, ncol=2, nlyr=5)
    wl.prop[1, 1, 2, 1] = 1.5  # ssalb > 1: c_check_inputs rejects
    ds = Disort(wl.options())
    with pytest.raises(RuntimeError, match="solves failed"):
        ds.forward(torch.from_numpy(wl.prop), {k: torch.from_numpy(v) for k, v in wl.bc.items()})
    st = ds.status().numpy()
    assert st[1, 1] == 1 and st.sum() == 1


def test_reduce_band(cuda_device):
    wl = W.make("amars_lw", nwave=16, ncol=3, nlyr=20)
    ds = Disort(wl.options())
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf))
    w = torch.rand(16, dtype=torch.float64)
    got = ds.reduce_band(flx, w).cpu()
    want = (flx.cpu() * w.view(-1, 1, 1, 1)).sum(0)
    assert torch.allclose(got, want, rtol=1e-13, atol=0)


import golden_lib as G  # noqa: E402


@pytest.mark.parametrize("name", G.NAMES)
def test_golden_fixtures(cuda_device, name):
    """Committed fixtures: extended-precision oracle, independent numpy solve, and (amars_sw_real) the real
    inputs of the reference's examples/amars_sw.cpp."""
    wl, ld, ind, valid = G.load(name)
    ds = Disort(wl.options())
    t = lambda a: None if a is None else torch.from_numpy(a).to(cuda_device)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)).cpu().numpy()
    assert G.relerr(flx, ld) < RTOL
    if ind is not None:
        assert G.relerr(flx, ind, valid) < RTOL
