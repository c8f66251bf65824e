"""Generates the opacity / interpolation / layer2level fixtures (tests/golden/opacity_*.npz) by running the
REFERENCE's own code (oracle/_ref/libharp_ref.so, built from /root/reference by oracle/ref_harp/Makefile)
on seeded inputs and on the reference's aerosol tables data/s8_k_fuller.txt, data/h2so4.txt.
Run from the repo root in the build container:  python tests/golden/make_golden_opacity.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_harp_lib as R  # noqa: E402
from pyharp_b200.opacity import read_table  # noqa: E402

REF = "/root/reference"


def interpn_cases(r):
    out = {}
    for ndim, nval in ((1, 1), (1, 2), (2, 2), (3, 1), (3, 3), (4, 2)):
        dims = [int(x) for x in r.integers(3, 9, ndim)]
        axes = [np.sort(r.uniform(-2, 5, d)) for d in dims]
        if ndim >= 2:
            axes[1] = axes[1][::-1].copy()  # a descending axis (locate handles both)
        data = r.normal(size=dims + [nval])
        npt = 400
        coord = np.stack([r.uniform(a.min() - 1.0, a.max() + 1.0, npt) for a in axes], 1)
        # exact nodes, end points and out-of-range points
        for k, a in enumerate(axes):
            coord[:len(a), k] = a
            coord[len(a), k] = a[0]
            coord[len(a) + 1, k] = a[-1]
        key = f"nd{ndim}_nv{nval}"
        out[key + "_coord"], out[key + "_data"] = coord, data
        for k, a in enumerate(axes):
            out[f"{key}_axis{k}"] = a
        out[key + "_out"] = R.interpn(coord, data, axes)
    return out


def main():
    r = np.random.default_rng(20261018 + 50)
    np.savez_compressed(os.path.join(HERE, "opacity_interpn.npz"), **interpn_cases(r))

    # aerosols exactly as examples/amars_sw.cpp:218-267 sets them up (species S8, H2SO4)
    weights = [256.0e-3, 98.0e-3]
    wn = np.linspace(2000.0, 50000.0, 64)
    conc = 10.0 ** r.uniform(-9, -5, (2, 6, 2))
    d = {"species_weights": np.array(weights), "wavenumber": wn, "conc": conc}
    for kind, fname, sid in (("s8_fuller", "s8_k_fuller.txt", 0), ("h2so4_simple", "h2so4.txt", 1)):
        path = os.path.join(REF, "data", fname)
        out, kwave, kdata = R.aerosol_forward(kind, path, weights, sid, wn, conc, by_wavenumber=True)
        d[kind + "_table"] = read_table(path)
        d[kind + "_out"], d[kind + "_kwave"], d[kind + "_kdata"] = out, kwave, kdata
        wl = np.concatenate([d[kind + "_table"][::7, 0], [0.01, 500.0]])  # table nodes + both out-of-range sides
        d[kind + "_wavelength"] = wl
        d[kind + "_out_wl"] = R.aerosol_forward(kind, path, weights, sid, wl, conc)[0]
    np.savez_compressed(os.path.join(HERE, "opacity_aerosol.npz"), **d)

    # RFM on a synthetic ck table (the reference's NetCDF tables are not shipped)
    nw, npres, nt = 12, 9, 5
    wave = np.linspace(1.0, 150.0, nw)
    lnp = np.log(np.logspace(5.5, 1.0, npres))
    dT = np.linspace(-40.0, 40.0, nt)
    kaxis = np.concatenate([wave, lnp, dT])
    kdata = r.uniform(-12.0, 3.0, (nw, npres, nt))
    krefatm = np.stack([lnp, np.linspace(290.0, 160.0, npres)])
    ncol, nlyr = 3, 8
    pres = np.exp(r.uniform(lnp.min() - 0.5, lnp.max() + 0.5, (ncol, nlyr)))
    temp = r.uniform(140.0, 320.0, (ncol, nlyr))
    conc = 10.0 ** r.uniform(-3, 1, (ncol, nlyr, 2))
    np.savez_compressed(os.path.join(HERE, "opacity_rfm.npz"), kaxis=kaxis, kdata=kdata, krefatm=krefatm, pres=pres,
                        temp=temp, conc=conc, out0=R.rfm_forward(kaxis, kdata, krefatm, pres, temp, conc, 0),
                        out1=R.rfm_forward(kaxis, kdata, krefatm, pres, temp, conc, 1))

    d = {}
    for nlyr in (1, 2, 3, 4, 5, 40):
        var = r.uniform(150.0, 300.0, (4, nlyr))
        d[f"var{nlyr}"] = var
        for order in (2, 4):
            for bl in (0, 1):
                for bu in (0, 1):
                    d[f"out{nlyr}_o{order}_l{bl}_u{bu}"] = R.layer2level(var, order, bl, bu)
    np.savez_compressed(os.path.join(HERE, "opacity_layer2level.npz"), **d)
    nrow, nx1 = 6, 21
    x1f = np.cumsum(r.uniform(500.0, 1500.0, nx1))
    rad = 3.39e6 + x1f
    area = np.tile(rad ** 2, (nrow, 1)) * r.uniform(0.99, 1.01, (nrow, 1))
    vol = area * np.gradient(x1f)[None, :]
    flx = r.uniform(10.0, 400.0, (nrow, nx1))
    np.savez_compressed(os.path.join(HERE, "opacity_spherical.npz"), flx=flx, x1f=x1f, area=area, vol=vol,
                        out=R.spherical_flux_correction(flx, x1f, area, vol))
    print("written:", [f for f in sorted(os.listdir(HERE)) if f.startswith("opacity_")])


if __name__ == "__main__":
    main()
