"""The device math of the CUDA kernels (pyharp_b200/csrc/disort_device.cuh) compiled for the host
by tests/hostsim, checked against the golden fixtures and the oracle without a GPU.  The product
never runs this way (no CPU path); this keeps the algorithm honest on a machine without a B200."""
import numpy as np
import pytest

import golden_lib as G
import hostsim_lib as H
import oracle_lib as O
from pyharp_b200 import workloads as W

RTOL = 1e-9


def _sim(wl):
    flx, st = H.flux(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                     wave_upper=wl.wave_upper)
    return flx, st


@pytest.mark.parametrize("name", G.NAMES)
def test_hostsim_layerwise_order(name):
    """The order of the warp-per-solve kernel (layer operators first, then one serial adding sweep; nstr <= 8)."""
    wl, ld, ind, valid = G.load(name)
    if wl.nstr > 8:
        pytest.skip("warp-per-solve form: nstr <= 8")
    flx, st = H.flux(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                     wave_upper=wl.wave_upper, layerwise=True)
    assert not st.any()
    G.check(flx, ld, RTOL)
    w = np.linspace(0.5, 1.5, wl.shape[0])
    bflx, st = H.flux(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                      wave_upper=wl.wave_upper, layerwise=True, weights=w)
    ref = (flx * w[:, None, None, None]).sum(0)
    assert np.max(np.abs(bflx - ref)) <= 1e-14 * np.max(np.abs(ref))


@pytest.mark.parametrize("name", G.NAMES)
def test_hostsim_vs_golden(name):
    wl, ld, ind, valid = G.load(name)
    flx, st = _sim(wl)
    assert not st.any()
    G.check(flx, ld, RTOL)  # fluxes and per-layer net-flux divergence
    if ind is not None:
        G.check(flx, ind, RTOL, valid)


@pytest.mark.parametrize("name,kw", [("test1a", {}), ("amars_sw", dict(nwave=40, ncol=2)),
                                     ("amars_lw", dict(nwave=16, ncol=2)), ("lbl", dict(nwave=300)),
                                     ("gcm", dict(nwave=2, ncol=8))])
def test_hostsim_vs_oracle(name, kw):
    wl = W.make(name, **kw)
    ref, _, st0 = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                            wave_upper=wl.wave_upper, nthreads=4)
    flx, st = _sim(wl)
    assert not st.any() and not st0.any()
    assert G.relerr(flx, ref) < RTOL


def test_layer_and_level_order():
    """pyharp is bottom-up (index 0 = surface), DISORT top-down: a beam through a pure absorber must give
    total-down = mu0 F exp(-tau_above/mu0) with tau_above accumulated from the LAST layer index."""
    nl = 5
    dtau = np.array([0.5, 0.4, 0.3, 0.2, 0.1])  # index 0 = surface layer
    prop = dtau.reshape(1, 1, nl, 1)
    bc = {"fbeam": np.array([[2.0]]), "umu0": np.array([[0.5]])}
    flx, st = H.flux(prop, bc, nstr=4, flags=O.LAMBER | O.ONLYFL)
    tau_above = np.concatenate([np.cumsum(dtau[::-1])[::-1], [0.0]])  # level i sits below layers i..nl-1
    np.testing.assert_allclose(flx[0, 0, :, 1], 0.5 * 2.0 * np.exp(-tau_above / 0.5), rtol=1e-13)
    np.testing.assert_allclose(flx[0, 0, :, 0], 0.0, atol=1e-15)


def test_energy_conservation():
    """omega -> 1, albedo = 1, no emission: net flux is constant with depth."""
    r = np.random.default_rng(11)
    nl = 20
    prop = np.zeros((3, 2, nl, 2))
    prop[..., 0] = 10.0 ** r.uniform(-2, 0.5, (3, 2, nl))
    prop[..., 1] = 1.0 - 1e-12
    bc = {"fbeam": np.full((3, 2), 5.0), "umu0": np.full((3, 2), 0.7), "albedo": np.ones((3, 2))}
    flx, st = H.flux(prop, bc, nstr=8, flags=O.LAMBER | O.ONLYFL)
    assert not st.any()
    net = flx[..., 1] - flx[..., 0]
    assert np.max(np.abs(net)) < 1e-6 * 5.0 * 0.7  # everything that enters leaves again


def test_layer_splitting_invariance():
    """Splitting homogeneous layers in two leaves the boundary fluxes unchanged (B.2 (iv))."""
    r = np.random.default_rng(12)
    nl = 6
    prop = np.zeros((2, 2, nl, 2 + 8))
    prop[..., 0] = 10.0 ** r.uniform(-2, 0.3, (2, 2, nl))
    prop[..., 1] = r.uniform(0.1, 0.99, (2, 2, nl))
    g = r.uniform(0.1, 0.8, (2, 2, nl))
    for k in range(8):
        prop[..., 2 + k] = g ** (k + 1)
    bc = {"fbeam": np.full((2, 2), 3.0), "umu0": np.full((2, 2), 0.45), "albedo": np.full((2, 2), 0.2)}
    a, _ = H.flux(prop, bc, nstr=8, flags=O.LAMBER | O.ONLYFL)
    split = np.repeat(prop, 2, axis=2)
    split[..., 0] *= 0.5
    b, _ = H.flux(split, bc, nstr=8, flags=O.LAMBER | O.ONLYFL)
    np.testing.assert_allclose(b[:, :, ::2], a, rtol=1e-11, atol=1e-13)


def test_linearity_in_sources():
    wl = W.make("amars_sw", nwave=6, ncol=2, nlyr=12)
    a, _ = _sim(wl)
    wl.bc["fbeam"] = 3.0 * wl.bc["fbeam"]
    b, _ = _sim(wl)
    np.testing.assert_allclose(b, 3.0 * a, rtol=1e-12, atol=1e-300)


def test_isothermal_equilibrium():
    """Thick isothermal column over a black surface at the same temperature: up = down = pi B in the interior."""
    nl = 30
    prop = np.full((1, 1, nl, 1), 3.0)
    temf = np.full((1, nl + 1), 280.0)
    bc = {"btemp": np.array([[280.0]]), "albedo": np.array([[0.0]])}
    flx, st = H.flux(prop, bc, temf, nstr=8, flags=O.LAMBER | O.ONLYFL | O.PLANCK, wave_lower=np.array([400.0]),
                     wave_upper=np.array([800.0]))
    piB = np.pi * O.lib().orc_plkavg(400.0, 800.0, 280.0)
    np.testing.assert_allclose(flx[0, 0, 5:20], piB, rtol=1e-10)


def test_status_codes():
    wl = W.make("amars_sw", nwave=3, ncol=2, nlyr=4)
    wl.prop[0, 1, 2, 0] = -1.0                      # negative optical depth
    wl.prop[2, 0, 1, 1] = 1.2                       # ssalb > 1
    wl.bc["albedo"][1, 1] = 1.5
    flx, st = _sim(wl)
    ref, _, st0 = O.forward(wl.prop, wl.bc, nstr=wl.nstr, flags=wl.flags)
    assert (st != 0).tolist() == (st0 != 0).tolist()
    assert st[0, 1] == 1 and st[2, 0] == 1 and st[1, 1] == 1 and st.sum() == 3
    assert np.isnan(flx[0, 1]).all() and np.isfinite(flx[0, 0]).all()
    # beam coinciding with a quadrature angle is rejected (DISORT: fatal input error)
    wl = W.make("amars_sw", nwave=1, ncol=1, nlyr=4)
    x, _ = np.polynomial.legendre.leggauss(4)
    wl.bc["umu0"][:] = 0.5 * (x[2] + 1)
    _, st = _sim(wl)
    assert st[0, 0] == 5


def _edge_problem(nstr, nl, ssa=None, zero_layers=(), thermal=False, seed=5):
    rng = np.random.default_rng(seed)
    nw, nc = 2, 3
    prop = np.zeros((nw, nc, nl, 2 + nstr))
    prop[..., 0] = 10 ** rng.uniform(-1, 0.5, (nw, nc, nl))
    prop[..., 1] = rng.uniform(0, 0.99, (nw, nc, nl)) if ssa is None else ssa
    g = rng.uniform(0, 0.8, (nw, nc, nl))
    for k in range(nstr):
        prop[..., 2 + k] = g ** (k + 1)
    for z in zero_layers:
        prop[:, :, z, 0] = 0.0
    bc = {"fbeam": np.full((nw, nc), 2.0), "umu0": np.full((nw, nc), 0.62), "albedo": np.full((nw, nc), 0.3)}
    flags, temf, kw = O.LAMBER | O.ONLYFL, None, {}
    if thermal:
        flags |= O.PLANCK
        temf = np.linspace(300, 200, nl + 1)[None, :].repeat(nc, 0)
        bc["btemp"] = np.full((nw, nc), 300.0)
        kw = dict(wave_lower=np.full(nw, 500.0), wave_upper=np.full(nw, 700.0))
    return prop, bc, temf, flags, kw


@pytest.mark.parametrize("thermal", [False, True])
def test_layers_of_zero_optical_depth(thermal):
    prop, bc, temf, flags, kw = _edge_problem(8, 6, zero_layers=(0, 3, 5), thermal=thermal)
    ref, _, st = O.forward(prop, bc, temf, nstr=8, flags=flags, **kw)
    got, st2 = H.flux(prop, bc, temf, nstr=8, flags=flags, **kw)
    assert not st.any() and not st2.any()
    assert G.relerr(got, ref) < RTOL
    # a layer without optical depth changes nothing: the fluxes on its two sides are equal
    np.testing.assert_allclose(got[:, :, 0], got[:, :, 1], rtol=1e-13)


def test_conservative_scattering_is_dithered():
    """ssalb = 1 exactly: DISORT replaces it by 1 - 1e-10 (a zero eigenvalue otherwise).  In that regime the
    eigenvalue k ~ 1e-5 makes every FP64 formulation lose digits: the FP64 oracle (banded LU, like cdisort) is
    only ~1e-7 from the extended-precision solve, the adding formulation of the kernels ~1e-9.  Parity is therefore
    checked against the extended-precision oracle at 1e-8."""
    prop, bc, temf, flags, kw = _edge_problem(8, 5, ssa=1.0)
    ld, _, st = O.forward(prop, bc, temf, nstr=8, flags=flags, long_double=True)
    f64, _, _ = O.forward(prop, bc, temf, nstr=8, flags=flags)
    got, st2 = H.flux(prop, bc, temf, nstr=8, flags=flags)
    assert not st.any() and not st2.any()
    assert G.relerr(got, ld) < 1e-8
    assert G.relerr(f64, ld) < 1e-5
    assert np.isfinite(got).all() and (got >= -1e-12).all()


def test_varying_cutoff_depth_across_solves():
    """Every third solve is cut off part-way down (abstau >= 10, c_disort_set's ncut) while its neighbours are not:
    the workspace slice and the per-thread state are re-used from solve to solve and must not leak between them."""
    from pyharp_b200 import workloads as W
    wl = W.make("amars_sw", nwave=50, ncol=7, nlyr=10)
    prop = wl.prop.copy()
    flat = prop.reshape(-1, prop.shape[2], prop.shape[3])
    rng = np.random.default_rng(7)
    for k in range(0, flat.shape[0], 3):
        flat[k, :, 0] = rng.uniform(0.5, 6.0, size=flat.shape[1])
        flat[k, :, 1] = rng.uniform(0.0, 0.3, size=flat.shape[1])
    ref, _, st = O.forward(prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                           wave_upper=wl.wave_upper)
    assert not st.any()
    out = H.flux(prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                 wave_upper=wl.wave_upper)
    flx = out[0] if isinstance(out, tuple) else out
    assert int((ref[:, :, 0, 1] == 0).sum()) > 50  # the cut-off really happens
    scale = np.max(np.abs(ref), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(flx - ref) / scale)) < 1e-9
