"""Fused band integration of the flux kernel (BAND instantiation, SolveParams::bflx_t) on the host simulator:
sum_n w_n F_n accumulated in the kernel's flux epilogue equals the weighted sum of the per-spectral-point fluxes
(examples/amars_lw.cpp:88), also with boundary fields shared by all spectral points (bc_bcast)."""
import numpy as np
import pytest

import hostsim_lib as H
from pyharp_b200 import workloads as W


@pytest.mark.parametrize("name,kw", [("amars_lw", dict(nwave=6, ncol=3, nlyr=12)), ("amars_sw", dict(nwave=5, ncol=2, nlyr=10)),
                                     ("lbl", dict(nwave=9, ncol=1, nlyr=8)), ("gcm", dict(nwave=3, ncol=2, nlyr=6))])
def test_band_accumulation_matches_weighted_sum(name, kw):
    wl = W.make(name, **kw)
    a = dict(nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower, wave_upper=wl.wave_upper)
    flx, st = H.flux(wl.prop, wl.bc, wl.temf, **a)
    assert not st.any()
    w = np.random.default_rng(3).uniform(0.1, 1.0, wl.shape[0])
    bflx, st = H.flux(wl.prop, wl.bc, wl.temf, weights=w, **a)
    assert not st.any() and bflx.shape == flx.shape[1:]
    ref = (flx * w[:, None, None, None]).sum(0)
    assert np.max(np.abs(bflx - ref)) <= 4e-16 * np.max(np.abs(ref)) * wl.shape[0]


def test_band_with_cutoff_levels():
    """Levels below the absorption cut-off add nothing (the non-band kernel writes zeros there)."""
    wl = W.make("amars_sw", nwave=4, ncol=2, nlyr=12)
    wl.prop[1, :, :6, 0] = 8.0   # thick absorbing lower half for one spectral point
    wl.prop[1, :, :6, 1] = 0.1
    a = dict(nstr=wl.nstr, flags=wl.flags)
    flx, st = H.flux(wl.prop, wl.bc, **a)
    assert (flx[1, :, 0, :] == 0.0).all()
    w = np.array([0.3, 0.2, 0.4, 0.1])
    bflx, _ = H.flux(wl.prop, wl.bc, weights=w, **a)
    ref = (flx * w[:, None, None, None]).sum(0)
    assert np.max(np.abs(bflx - ref)) <= 1e-15 * np.max(np.abs(ref))


def test_band_boundary_fields_per_column():
    wl = W.make("amars_lw", nwave=5, ncol=4, nlyr=9, variant=1)
    a = dict(nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower, wave_upper=wl.wave_upper)
    w = np.linspace(0.1, 0.5, 5)
    full, _ = H.flux(wl.prop, wl.bc, wl.temf, weights=w, **a)
    per_col = {k: v[0] for k, v in wl.bc.items()}  # identical for every spectral point in this workload
    got, st = H.flux(wl.prop, per_col, wl.temf, weights=w, bc_bcast=True, **a)
    assert not st.any() and np.array_equal(got, full)
