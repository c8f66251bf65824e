"""Seeded inputs for the general (user levels / user angles / azimuth series / NT correction) path,
shared by the CPU host-simulator tests and the GPU parity tests."""
import numpy as np

import oracle_lib as O

# (nstr, nlyr, nmom, beam, thermal, usrang, usrtau, intcor)
CASES = [(8, 5, 8, 1, 0, 1, 0, 0), (8, 5, 8, 1, 0, 1, 1, 1), (8, 5, 12, 1, 0, 1, 1, 1), (4, 7, 4, 1, 1, 1, 1, 0),
         (16, 3, 16, 1, 1, 1, 0, 1), (8, 6, 8, 0, 1, 1, 1, 0), (8, 6, 8, 1, 0, 0, 0, 0), (8, 6, 8, 1, 1, 0, 1, 0),
         (6, 4, 10, 1, 0, 1, 0, 1), (32, 3, 32, 1, 1, 1, 1, 1), (8, 40, 8, 1, 1, 1, 0, 1)]


def make(case, nw=2, nc=2):
    nstr, L, nmom, beam, thermal, usrang, usrtau, intcor = case
    rng = np.random.default_rng(1000 + sum(case) * 7 + L)
    prop = np.zeros((nw, nc, L, 2 + nmom))
    prop[..., 0] = 10 ** rng.uniform(-2, 0.3, (nw, nc, L))
    prop[..., 1] = rng.uniform(0.1, 0.99, (nw, nc, L))
    g = rng.uniform(0.2, 0.85, (nw, nc, L))
    for k in range(nmom):
        prop[..., 2 + k] = g ** (k + 1)
    bc = {"albedo": rng.uniform(0, 0.8, (nw, nc))}
    flags, kw, temf = O.LAMBER, {}, None
    if beam:
        bc.update(fbeam=rng.uniform(1, 5, (nw, nc)), umu0=rng.uniform(0.2, 0.9, (nw, nc)), phi0=np.full((nw, nc), 20.0))
    if thermal:
        flags |= O.PLANCK
        temf = np.linspace(300, 200, L + 1)[None, :].repeat(nc, 0)
        bc.update(btemp=np.full((nw, nc), 305.0), ttemp=np.full((nw, nc), 150.0), temis=np.full((nw, nc), 0.7))
        kw.update(wave_lower=np.full(nw, 300.0), wave_upper=np.full(nw, 800.0))
    if usrang:
        flags |= O.USRANG
        kw["umu"] = [-1, -0.6, -0.2, 0.15, 0.55, 1.0]
    tot = prop[..., 0].sum(-1).min()
    if usrtau:
        flags |= O.USRTAU
        kw["utau"] = [0, 0.3 * tot, 0.7 * tot, tot]
    if intcor:
        flags |= O.INTCOR | O.OLDINTCOR
    kw["phi"] = [0.0, 60.0, 150.0]
    return dict(prop=prop, bc=bc, temf=temf, nstr=nstr, nmom=nmom, flags=flags, accur=0.0, **kw)


def flag_tokens(flags):
    names = {O.USRTAU: "usrtau", O.USRANG: "usrang", O.LAMBER: "lamber", O.PLANCK: "planck", O.ONLYFL: "onlyfl",
             O.INTCOR: "intensity_correction", O.OLDINTCOR: "old_intensity_correction"}
    return ",".join(v for k, v in names.items() if flags & k)
