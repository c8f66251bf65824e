"""Wide-dynamic-range flux problems for the kernels' fast arithmetic (branch-free rcp / rsqrt seeds, half-angle
Jacobi rotations, integer-pipe pivot checks): optical depths from 1e-12 (1e-7 with thermal emission) to 1e3 per layer, single-scattering albedos
from 0 to 1 - 1e-7, Henyey-Greenstein phase functions up to g = 0.95, grazing to overhead sun, 40 K to 1000 K,
beam and thermal sources together, every compiled stream count.  Test infrastructure only."""
import numpy as np


def make(nstr, seed, nwave=6, ncol=40, nlyr=17, thermal=True, beam=True, nprop=None):
    r = np.random.default_rng(1000 * nstr + seed)
    nprop = (2 + nstr) if nprop is None else nprop
    prop = np.zeros((nwave, ncol, nlyr, nprop))
    # (with thermal emission the thin end stops at 1e-7: DISORT's linear-in-tau Planck source divides the Planck
    # difference across a layer by its optical depth, and the double-precision reference itself loses its digits
    # below that -- checked against the 80-bit build of the oracle, see tests/test_hostsim_golden.py)
    prop[..., 0] = 10.0 ** r.uniform(-7.0 if thermal else -12.0, 3.0, (nwave, ncol, nlyr))
    # a few exactly transparent layers
    if not thermal:
        prop[..., 0] = np.where(r.uniform(size=prop[..., 0].shape) < 0.03, 0.0, prop[..., 0])
    if nprop > 1:
        ssa = 1.0 - 10.0 ** r.uniform(-7.0, 0.0, (nwave, ncol, nlyr))
        ssa = np.where(r.uniform(size=ssa.shape) < 0.1, 0.0, ssa)   # pure absorbers in between
        prop[..., 1] = ssa
    if nprop > 2:
        g = r.uniform(0.0, 0.95, (nwave, ncol, nlyr))
        for k in range(nprop - 2):
            prop[..., 2 + k] = g ** (k + 1)
    bc = {"albedo": r.uniform(0.0, 1.0, (nwave, ncol))}
    bc["albedo"][0, :3] = (0.0, 1.0, 0.5)
    flags = "lamber,quiet,onlyfl"
    temf = wlo = whi = None
    if beam:
        mu0 = r.uniform(0.02, 1.0, (nwave, ncol))
        x, _ = np.polynomial.legendre.leggauss(nstr // 2)
        for q in 0.5 * (x + 1.0):   # DISORT rejects a beam on a quadrature angle
            mu0 = np.where(np.abs(mu0 - q) / mu0 < 1e-3, mu0 * 1.003, mu0)
        bc["umu0"] = np.minimum(mu0, 1.0)
        bc["fbeam"] = 10.0 ** r.uniform(-3.0, 3.0, (nwave, ncol))
    if thermal:
        flags += ",planck"
        # smooth profiles (a few kelvin per level) anywhere between 40 K and 1000 K
        temf = r.uniform(60.0, 980.0, (ncol, 1)) + np.cumsum(r.uniform(-4.0, 4.0, (ncol, nlyr + 1)), axis=1)
        temf[0, :] = 250.0                      # an isothermal column
        bc["btemp"] = r.uniform(40.0, 1000.0, (nwave, ncol))
        bc["ttemp"] = r.uniform(0.0, 300.0, (nwave, ncol))
        bc["temis"] = r.uniform(0.0, 1.0, (nwave, ncol))
        wlo = 10.0 ** r.uniform(0.0, 3.3, nwave)
        whi = wlo * r.uniform(1.001, 3.0, nwave)
    return dict(prop=prop, bc=bc, temf=temf, nstr=nstr, flags=flags, wave_lower=wlo, wave_upper=whi)
