"""amars_sw as a radiation band (examples/amars_sw.cpp:218-301 as ONE band-level call): S8 + H2SO4 aerosol attenuators
(`S8Fuller`, `H2SO4Simple` on the reference's own optical tables, committed as the fixture tests/golden/opacity_aerosol.npz),
property mixing, nstr = 8 beam solve with a Lambertian surface, weighted band sum.  Used by bench.py for the end-to-end
number of BASELINE config 2 at the band level (only concentrations, per-point boundary fields and the band flux cross PCIe)."""
import os

import numpy as np

_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "opacity_aerosol.npz")


def make(nwave=500, ncol=1024, nlyr=40, seed=20261018 + 77, device=None):
    """-> (band, inputs): a pyharp_b200.RadiationBand and its host inputs dict(conc, dz, bc, weights, wavenumber)."""
    from .band import RadiationBand, RadiationBandOptions
    from .disort import DisortOptions
    from .opacity import AttenuatorOptions, H2SO4Simple, S8Fuller
    z = np.load(_FIXTURE)
    r = np.random.default_rng(seed)
    wn = np.linspace(2000.0, 50000.0, nwave)  # cm^-1, examples/amars_sw.cpp:200-208
    # aerosol columns: a haze whose density falls off with height (index 0 = surface), nothing in the top layers
    conc = np.abs(r.normal(1.0, 0.5, (ncol, nlyr, 2))) * 3e-3 * np.exp(-np.linspace(0.0, 5.0, nlyr))[None, :, None]
    conc[:, -3:, :] = 0.0
    dz = np.linspace(400.0, 1600.0, nlyr)
    # solar beam: Planck(5772 K) shape over the band (arbitrary scale), one zenith angle and albedo per column
    x = 1.438786 * wn / 5772.0
    fb = (wn ** 3 / np.expm1(x))
    fbeam = np.ascontiguousarray((fb / fb.max())[:, None] * r.uniform(0.8, 1.2, (1, ncol)))
    mu0 = r.uniform(0.2, 1.0, ncol)
    g, _ = np.polynomial.legendre.leggauss(4)
    for q in 0.5 * (g + 1.0):  # keep the beam off the quadrature angles (c_disort rejects |mu0 - mu_q| / mu0 < 1e-4)
        mu0 = np.where(np.abs(mu0 - q) / mu0 < 1e-3, mu0 * 1.003, mu0)
    bc = {"fbeam": fbeam, "umu0": np.ascontiguousarray(np.broadcast_to(np.minimum(mu0, 1.0), (nwave, ncol))),
          "albedo": np.ascontiguousarray(np.broadcast_to(r.uniform(0.0, 0.4, ncol), (nwave, ncol)))}
    w = r.uniform(0.5, 1.5, nwave)
    w /= w.sum()
    atts = {}
    for kind, cls, sid in (("s8_fuller", S8Fuller, 0), ("h2so4_simple", H2SO4Simple, 1)):
        op = AttenuatorOptions().species_weights(list(z["species_weights"])).species_names(["S8", "H2SO4"]).species_ids([sid])
        atts[kind] = cls(op, table=z[kind + "_table"])
    dop = DisortOptions().flags("lamber,quiet,onlyfl,intensity_correction,old_intensity_correction")
    dop.ds().nstr = dop.ds().nmom = 8
    bo = RadiationBandOptions().attenuators(atts).disort(dop).nlyr(nlyr).ncol(ncol)
    bo.wave_lower([0.0] * nwave).wave_upper([0.0] * nwave).weights(w)
    band = RadiationBand(bo, wave=wn, wave_is_wavenumber=True) if device is None else RadiationBand(bo, wave=wn, wave_is_wavenumber=True, device=device)
    return band, dict(conc=np.ascontiguousarray(conc), dz=dz, bc=bc, weights=w, wavenumber=wn)
