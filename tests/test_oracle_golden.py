"""Pins the CPU oracle: DISORT test problem 1a (the configuration tests/test_disort.cpp sets up),
the committed golden fixtures (extended-precision oracle + independent numpy solve) and unit
checks of its building blocks.  CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O


def test_test1a_published_values():
    t = json.load(open(os.path.join(G.GOLDEN_DIR, "test1a.json")))
    nstr = t["nstr"]
    prop = np.zeros((1, 1, 1, 2 + nstr))
    prop[..., 0], prop[..., 1] = t["dtau"], t["ssalb"]
    bc = {"umu0": np.array([[t["umu0"]]]), "fbeam": np.array([[np.pi * t["fbeam_over_pi"]]])}
    flags = O.USRTAU | O.USRANG | O.LAMBER | O.INTCOR | O.OLDINTCOR
    flx, uu, st = O.forward(prop, bc, nstr=nstr, flags=flags, utau=t["utau"], umu=t["umu"], phi=t["phi"])
    assert not st.any()
    # pyharp order: index 0 = bottom level
    fdn = np.array(t["rfldir"]) + np.array(t["rfldn"])
    np.testing.assert_allclose(flx[0, 0, ::-1, 1], fdn, rtol=1e-9)
    np.testing.assert_allclose(flx[0, 0, ::-1, 0], t["flup"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(uu[0, 0, 0], np.array(t["uu"]), rtol=1e-9, atol=1e-12)
    # and the 6-digit numbers printed in the DISORT distribution
    p = t["published_6_digits"]
    np.testing.assert_allclose(uu[0, 0, 0], np.array(p["uu"], dtype=float), rtol=5e-6, atol=1e-12)
    assert abs(flx[0, 0, 1, 0] / p["flup0"] - 1) < 5e-6


@pytest.mark.parametrize("name", G.NAMES)
def test_oracle_vs_golden(name):
    wl, ld, ind, valid = G.load(name)
    flx, _, st = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                           wave_upper=wl.wave_upper)
    assert not st.any()
    assert G.relerr(flx, ld) < 1e-10
    if ind is not None:
        assert G.relerr(flx, ind, valid) < 1e-10


def test_serial_and_threaded_oracle_agree():
    wl, ld, _, _ = G.load("amars_sw")
    a, _, _ = O.forward(wl.prop, wl.bc, nstr=wl.nstr, flags=wl.flags, nthreads=1)
    b, _, _ = O.forward(wl.prop, wl.bc, nstr=wl.nstr, flags=wl.flags, nthreads=4)
    assert np.array_equal(a, b)


def test_double_gauss_quadrature():
    for n in (1, 2, 4, 8, 16):
        mu, w = np.zeros(n), np.zeros(n)
        dp = C.POINTER(C.c_double)
        O.lib().orc_qgausn(n, mu.ctypes.data_as(dp), w.ctypes.data_as(dp))
        x, ww = np.polynomial.legendre.leggauss(n)
        np.testing.assert_allclose(np.sort(mu), 0.5 * (x + 1), rtol=1e-14)
        np.testing.assert_allclose(w[np.argsort(mu)], 0.5 * ww, rtol=1e-13)


def test_plkavg_known_answers():
    pk = O.lib().orc_plkavg
    sigma = 5.67032e-8  # the legacy constant DISORT uses
    for T in (150.0, 300.0, 1000.0):
        # whole spectrum: sigma T^4 / pi
        assert abs(pk(0.0, 1.0e6, T) / (sigma * T ** 4 / np.pi) - 1) < 1e-5
        # additivity over adjacent bands and the narrow-band Simpson branch against numeric quadrature
        assert abs((pk(100., 600., T) + pk(600., 2000., T)) / pk(100., 2000., T) - 1) < 2e-6
        lo, hi = 999.0, 1000.0
        v = np.linspace(lo, hi, 2001)
        c1, c2 = 1.1910429e-8, 1.438786  # W m^-2 sr^-1 (cm^-1)^-4 ; cm K
        ref = np.trapezoid(c1 * v ** 3 / np.expm1(c2 * v / T), v)
        assert abs(pk(lo, hi, T) / ref - 1) < 1e-4
    assert pk(10.0, 20.0, 0.0) == 0.0


def test_asymtx_against_numpy():
    r = np.random.default_rng(5)
    dp = C.POINTER(C.c_double)
    for n in (2, 4, 8):
        # a matrix with real positive spectrum, like DISORT's (alpha-beta)(alpha+beta)
        a = r.uniform(0.1, 1.0, (n, n))
        s = a @ a.T + n * np.eye(n)
        d = np.diag(r.uniform(0.5, 2.0, n))
        m = np.linalg.inv(d) @ s @ d
        evec, ev = np.zeros((n, n)), np.zeros(n)
        rc = O.lib().orc_asymtx(np.ascontiguousarray(m).ctypes.data_as(dp), n, evec.ctypes.data_as(dp),
                                ev.ctypes.data_as(dp))
        assert rc == 0
        np.testing.assert_allclose(np.sort(ev), np.sort(np.linalg.eigvals(m).real), rtol=1e-12)
