"""Parity of the CUDA path (through the C ABI) against the CPU oracle: relative tolerance 1e-9
on fluxes (BASELINE.json north_star), measured against the largest flux of the solve."""
import numpy as np
import pytest
import torch

import golden_lib as G
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


FORMS = ["thread_per_solve", "warp_per_solve", "layer_split"]


def _skip_form(form, nstr):
    if form == "warp_per_solve" and nstr > 8:
        pytest.skip("the warp-per-solve form is instantiated for nstr <= 8")
    if form == "layer_split" and nstr not in (4, 8, 12, 16, 24, 32):
        pytest.skip("the layer-split form is instantiated for nstr = 4, 8 and nstr >= 12")


def _with_form(ds, form):
    """thread_per_solve = the fused kernel; warp_per_solve (nstr <= 8) / layer_split (nstr >= 12) forced on."""
    ds.set_warp_solve_max(1 << 30 if form == "warp_per_solve" else 0)
    return ds.set_layer_split(1 if form == "layer_split" else 0)


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
@pytest.mark.parametrize("form", FORMS)
def test_flux_parity(cuda_device, name, kw, where, form):
    """Every form of the flux kernel: small batches default to the warp-per-solve form (nstr <= 8) and nstr >= 12 to the
    layer-split form; each is forced off / on here so that each case runs on each form that exists for its stream count."""
    wl = W.make(name, **kw)
    _skip_form(form, wl.nstr)
    ref = _oracle(wl)
    ds = _with_form(Disort(wl.options()), form)
    dev = "cpu" if where == "host" else cuda_device
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf))
    assert flx.device.type == ("cpu" if where == "host" else "cuda")
    assert flx.shape == ref.shape
    err = _relerr(flx.cpu().numpy(), ref)
    assert err < RTOL, err
    assert G.diverr(flx.cpu().numpy(), ref) < RTOL  # net-flux divergence per layer (what the heating rate consumes)
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
@pytest.mark.parametrize("form", FORMS)
def test_golden_fixtures(cuda_device, name, form):
    """Committed fixtures: extended-precision oracle, independent numpy solve, and (amars_sw_real) the real
    inputs of the reference's examples/amars_sw.cpp; on every form of the flux kernel."""
    wl, ld, ind, valid = G.load(name)
    _skip_form(form, wl.nstr)
    ds = _with_form(Disort(wl.options()), form)
    t = lambda a: None if a is None else torch.from_numpy(a).to(cuda_device)
    flx = ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)).cpu().numpy()
    G.check(flx, ld, RTOL)  # fluxes and per-layer net-flux divergence
    if ind is not None:
        G.check(flx, ind, RTOL, valid)


@pytest.mark.parametrize("nstr", [2, 4, 6, 8, 12, 16, 24, 32])
@pytest.mark.parametrize("thermal", [False, True])
def test_every_stream_count(cuda_device, nstr, thermal):
    """All K1 instantiations (nn = 1, 2, 3, 4, 6, 8, 12, 16), beam + Lambertian surface, optionally thermal, with
    Henyey-Greenstein phase functions truncated by delta-M."""
    rng = np.random.default_rng(100 + nstr + int(thermal))
    nw, nc, nl = 3, 5, 9
    prop = np.zeros((nw, nc, nl, 2 + nstr))
    prop[..., 0] = 10.0 ** rng.uniform(-3, 0.4, (nw, nc, nl))
    prop[..., 1] = rng.uniform(0.0, 0.999, (nw, nc, nl))
    g = rng.uniform(0.0, 0.85, (nw, nc, nl))
    for k in range(nstr):
        prop[..., 2 + k] = g ** (k + 1)
    bc = {"fbeam": rng.uniform(1, 4, (nw, nc)), "umu0": rng.uniform(0.15, 0.95, (nw, nc)), "albedo": rng.uniform(0, 1, (nw, nc))}
    x, _ = np.polynomial.legendre.leggauss(nstr // 2)
    for q in 0.5 * (x + 1):  # keep the beam off the quadrature angles
        bc["umu0"] = np.where(np.abs(bc["umu0"] - q) / bc["umu0"] < 1e-3, bc["umu0"] * 1.01, bc["umu0"])
    tokens, temf, kw = "lamber,quiet,onlyfl", None, {}
    if thermal:
        tokens += ",planck"
        temf = np.linspace(290.0, 180.0, nl + 1)[None, :].repeat(nc, 0)
        bc.update(btemp=np.full((nw, nc), 295.0), ttemp=np.full((nw, nc), 120.0), temis=np.full((nw, nc), 0.3))
        kw = dict(wave_lower=np.array([100.0, 600.0, 1500.0]), wave_upper=np.array([600.0, 1500.0, 1510.0]))
    wl = W.Workload("streams", nstr, tokens, prop, bc, temf, kw.get("wave_lower"), kw.get("wave_upper"))
    ref = _oracle(wl)
    t = lambda a: None if a is None else torch.from_numpy(a).to(cuda_device)
    # the fused kernel, then (where they exist) the warp-per-solve / layer-split forms
    for form in ("thread_per_solve", "warp_per_solve", "layer_split"):
        if (form == "layer_split" and nstr not in (4, 8, 12, 16, 24, 32)) or (form == "warp_per_solve" and nstr > 8):
            continue
        ds = _with_form(Disort(wl.options()), form)
        flx = ds.forward(t(prop), {k: t(v) for k, v in bc.items()}, t(temf)).cpu().numpy()
        assert _relerr(flx, ref) < RTOL
        assert G.diverr(flx, ref) < RTOL


@pytest.mark.parametrize("nstr", [4, 8, 12, 16])
def test_layer_split_matches_fused_kernel(cuda_device, nstr):
    """nstr >= 12: the layer-split form against the fused kernel on a batch with absorption cut-offs at different
    levels, zero-thickness layers and a rejected input: same status words,
    fluxes equal to rounding (both execute the same statements; only the launch structure differs)."""
    rng = np.random.default_rng(7 + nstr)
    nw, nc, nl = 5, 77, 23
    prop = np.zeros((nw, nc, nl, 2 + nstr))
    prop[..., 0] = 10.0 ** rng.uniform(-3, 0.9, (nw, nc, nl))
    prop[..., 1] = rng.uniform(0.0, 0.999, (nw, nc, nl))
    prop[:, ::7, :, 1] *= 0.05                     # strongly absorbing columns: cut off part of the way down
    prop[:, 3::11, 5:8, 0] = 0.0                   # zero-thickness layers
    g = rng.uniform(0.0, 0.85, (nw, nc, nl))
    for k in range(nstr):
        prop[..., 2 + k] = g ** (k + 1)
    prop[2, 40, 9, 1] = 1.25                       # c_check_inputs: ssalb > 1
    bc = {"fbeam": rng.uniform(1, 4, (nw, nc)), "umu0": rng.uniform(0.15, 0.95, (nw, nc)), "albedo": rng.uniform(0, 1, (nw, nc))}
    x, _ = np.polynomial.legendre.leggauss(nstr // 2)
    for q in 0.5 * (x + 1):
        bc["umu0"] = np.where(np.abs(bc["umu0"] - q) / bc["umu0"] < 1e-3, bc["umu0"] * 1.01, bc["umu0"])
    wl = W.Workload("split", nstr, "lamber,quiet,onlyfl", prop, bc, None, None, None)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    status = []
    for split in (0, 1):
        ds = Disort(wl.options()).set_layer_split(split).set_warp_solve_max(0)
        with pytest.raises(RuntimeError, match="solves failed"):
            ds.forward(t(prop), {k: t(v) for k, v in bc.items()})
        status.append(ds.status().cpu().numpy())
    assert np.array_equal(status[0], status[1]) and status[0][2, 40] == 1 and status[0].sum() == 1
    # fluxes of the accepted solves: drop the rejected input and compare the two forms
    prop[2, 40, 9, 1] = 0.5
    res = []
    for split in (0, 1):
        ds = Disort(wl.options()).set_layer_split(split).set_warp_solve_max(0)
        res.append(ds.forward(t(prop), {k: t(v) for k, v in bc.items()}).cpu().numpy())
        assert int(ds.status().abs().sum()) == 0
    scale = np.max(np.abs(res[0]), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(res[0] - res[1]) / scale)) < 1e-12
    assert (res[0][:, ::7, 0, 1] == 0).any()       # some columns are cut off before the surface
    assert _relerr(res[1], _oracle(wl)) < RTOL


def test_layer_split_random_batch_thermal_and_beam(cuda_device):
    """nstr = 16, 3 x 1500 solves with beam AND thermal sources, random optical depths over five decades: the team adding
    kernel (groups of 8 solves per warp, the last group partly filled) against the fused kernel, and a sample against the
    oracle."""
    nstr, nw, nc, nl = 16, 3, 1500 + 3, 17
    rng = np.random.default_rng(2026)
    prop = np.zeros((nw, nc, nl, 2 + nstr))
    prop[..., 0] = 10.0 ** rng.uniform(-4, 1.0, (nw, nc, nl))
    prop[..., 1] = rng.uniform(0.0, 0.9999, (nw, nc, nl))
    g = rng.uniform(0.0, 0.9, (nw, nc, nl))
    for k in range(nstr):
        prop[..., 2 + k] = g ** (k + 1)
    bc = {"fbeam": rng.uniform(0.5, 4, (nw, nc)), "umu0": rng.uniform(0.1, 0.99, (nw, nc)), "albedo": rng.uniform(0, 1, (nw, nc)),
          "btemp": np.full((nw, nc), 290.0), "ttemp": np.full((nw, nc), 100.0), "temis": np.full((nw, nc), 0.5)}
    x, _ = np.polynomial.legendre.leggauss(nstr // 2)
    for q in 0.5 * (x + 1):
        bc["umu0"] = np.where(np.abs(bc["umu0"] - q) / bc["umu0"] < 1e-3, bc["umu0"] * 1.01, bc["umu0"])
    temf = np.linspace(280.0, 190.0, nl + 1)[None, :].repeat(nc, 0) + rng.uniform(-5, 5, (nc, 1))
    wl = W.Workload("split_rand", nstr, "lamber,quiet,onlyfl,planck", prop, bc, temf, np.array([200.0, 700.0, 1400.0]),
                    np.array([700.0, 1400.0, 1420.0]))
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    res = []
    for split in (0, 1):
        ds = Disort(wl.options()).set_layer_split(split)
        res.append(ds.forward(t(prop), {k: t(v) for k, v in bc.items()}, t(temf)).cpu().numpy())
        assert int(ds.status().abs().sum()) == 0
    scale = np.max(np.abs(res[0]), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(res[0] - res[1]) / scale)) < 1e-11
    sub = W.Workload("split_rand_sub", nstr, wl.flag_tokens, prop[:, :40], {k: v[:, :40] for k, v in bc.items()}, temf[:40],
                     wl.wave_lower, wl.wave_upper)
    assert _relerr(res[1][:, :40], _oracle(sub)) < RTOL
