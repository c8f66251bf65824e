"""Seeded synthetic atmospheres of the shapes BASELINE.json names.

Shared by the parity tests, smoke() and bench.py so the oracle, the CUDA path and the CPU
timing all see identical inputs.  numpy only; nothing here computes radiative transfer.

  test1a    tests/test_disort.cpp:16-47 literally (DISORT test problem 1a), flux part
  amars_sw  examples/amars_sw.cpp:43-65,262-278 shape: nstr=8, nprop=2 (tau, ssa), beam, albedo=1
  amars_lw  examples/amars_lw.cpp:18-38,62-80 shape: nstr=8, nprop=1, planck [1,150] cm^-1, btemp
  lbl       examples/amarsw-lbl.yaml shape: nstr=4, nprop=2, line-like tau, beam mu0=0.6
  gcm       BASELINE.json configs[4]: nstr=16, nprop=18 (tau, ssa, 16 HG moments), aerosol layers
"""
from dataclasses import dataclass, field

import numpy as np

SEED = 20261018
_CONFIG_INDEX = {"test1a": 0, "amars_sw": 1, "amars_lw": 2, "lbl": 3, "gcm": 4}

USRTAU, USRANG, LAMBER, PLANCK, ONLYFL = 1, 2, 4, 8, 16


@dataclass
class Workload:
    name: str
    nstr: int
    flag_tokens: str
    prop: np.ndarray                      # (nwave, ncol, nlyr, nprop), layer 0 = surface
    bc: dict = field(default_factory=dict)  # name -> (nwave, ncol)
    temf: np.ndarray = None               # (ncol, nlyr+1), level 0 = surface
    wave_lower: np.ndarray = None
    wave_upper: np.ndarray = None

    @property
    def shape(self):
        return self.prop.shape

    @property
    def nsolve(self):
        return self.prop.shape[0] * self.prop.shape[1]

    @property
    def flags(self):
        bits = {"usrtau": USRTAU, "usrang": USRANG, "lamber": LAMBER, "planck": PLANCK, "onlyfl": ONLYFL}
        out = 0
        for t in self.flag_tokens.split(","):
            out |= bits.get(t.strip(), 0)
        return out

    @property
    def beam(self):
        return "fbeam" in self.bc

    @property
    def thermal(self):
        return bool(self.flags & PLANCK)

    def options(self):
        from .disort import DisortOptions
        nwave, ncol, nlyr, _ = self.prop.shape
        op = DisortOptions().flags(self.flag_tokens).nwave(nwave).ncol(ncol)
        if self.wave_lower is not None:
            op.wave_lower(list(self.wave_lower)).wave_upper(list(self.wave_upper))
        d = op.ds()
        d.nlyr, d.nstr, d.nmom = nlyr, self.nstr, self.nstr
        d.nphi = d.ntau = d.numu = 1
        return op

    def flops_per_solve(self):
        """Algorithmic FP64 flops W of the reference algorithm (SURVEY.md section 8d / BASELINE.md section 3)."""
        _, _, L, _ = self.prop.shape
        N, n = self.nstr, self.nstr // 2
        kl = 3 * n - 1
        beam, thermal = int(self.beam), int(self.thermal)
        w = L * (N ** 3 + 29 * n ** 3 + beam * (2.0 / 3.0 * N ** 3 + 2 * N ** 2))
        w += 2 * N * L * kl * (kl + kl) + 2 * N * L * (2 * kl + kl)
        w += thermal * L * (2.0 / 3.0 * N ** 3 + 4 * N ** 2) + (L + 1) * (2 * N ** 2 + 4 * N)
        return float(w)

    def bytes_per_solve(self):
        """Algorithmic HBM bytes Q: inputs once, fluxes once, int32 status."""
        _, _, L, nprop = self.prop.shape
        return 8 * (L * nprop + len(self.bc) + int(self.thermal) * (L + 1) + 2 * (L + 1)) + 4


def _rng(name, extra=0):
    return np.random.default_rng(SEED + _CONFIG_INDEX[name] + 1000 * extra)


def _planck_wn(wn_cm, T):
    """Planck spectral radiance per wavenumber (W m^-2 sr^-1 / cm^-1), for the solar beam shape."""
    c1, c2 = 1.191042e-8, 1.4387769
    return c1 * wn_cm ** 3 / np.expm1(c2 * wn_cm / T)


def make(name, nwave=None, ncol=None, nlyr=None, variant=0):
    r = _rng(name, variant)
    if name == "test1a":
        nwave, ncol, nlyr, nstr = nwave or 10, ncol or 1, 1, 16
        prop = np.zeros((nwave, ncol, nlyr, 2 + nstr))
        prop[..., 0] = 0.03125
        prop[..., 1] = 0.2
        bc = {"umu0": np.full((nwave, ncol), 0.1), "fbeam": np.full((nwave, ncol), np.pi / 0.1)}
        return Workload(name, nstr, "lamber,quiet,onlyfl", prop, bc)

    if name == "amars_sw":
        nwave, ncol, nlyr, nstr = nwave or 500, ncol or 1, nlyr or 40, 8
        wn = np.linspace(2000.0, 50000.0, nwave)
        # aerosol column: extinction falls off with height (index 0 = surface), spectral slope per
        # wavenumber, ssa between 1e-5 and 0.9999 like the S8 / H2SO4 tables
        z = np.linspace(0.0, 1.0, nlyr)[None, None, :]
        colfac = 10.0 ** r.uniform(-1.0, 0.5, (1, ncol, 1))
        spec = 10.0 ** r.uniform(-3.0, 0.3, (nwave, 1, 1))
        tau = spec * colfac * np.exp(-6.0 * z) * 10.0 ** r.uniform(-0.3, 0.3, (nwave, ncol, nlyr))
        tau = np.maximum(tau, 1e-8)
        ssa = np.clip(1.0 - 10.0 ** r.uniform(-4.0, 0.0, (nwave, 1, 1)) * r.uniform(0.5, 1.0, (nwave, ncol, nlyr)),
                      1e-5, 0.9999)
        prop = np.stack([tau, ssa], axis=-1)
        fbeam = np.pi * _planck_wn(wn, 5772.0) * 2.92842e-5 * 0.7 * (wn[1] - wn[0] if nwave > 1 else 1.0)
        bc = {"fbeam": np.repeat(fbeam[:, None], ncol, 1), "umu0": np.ones((nwave, ncol)),
              "albedo": np.ones((nwave, ncol))}
        return Workload(name, nstr, "lamber,quiet,onlyfl,intensity_correction,old_intensity_correction", prop, bc)

    if name == "amars_lw":
        nwave, ncol, nlyr, nstr = nwave or 32, ncol or 1, nlyr or 80, 8
        # g-points sorted by opacity; per-layer tau log-uniform 1e-4 .. 1e1 across g
        g = np.sort(r.uniform(0.0, 1.0, nwave))[:, None, None]
        tau = 10.0 ** (-4.0 + 5.0 * g + r.uniform(-0.5, 0.5, (nwave, ncol, nlyr)) * 0.6)
        prop = tau[..., None]
        if variant == 0:
            temf = np.repeat(np.linspace(300.0, 150.0, nlyr + 1)[None, :], ncol, 0)
            temf = temf + r.uniform(-5.0, 5.0, (ncol, 1))
        else:
            temf = np.full((ncol, nlyr + 1), 300.0)
        alb = 1.0 if variant == 0 else 0.0
        bc = {"albedo": np.full((nwave, ncol), alb), "btemp": np.repeat(temf[None, :, 0], nwave, 0)}
        return Workload(name, nstr, "lamber,quiet,onlyfl,planck", prop, bc, temf,
                        np.full(nwave, 1.0), np.full(nwave, 150.0))

    if name == "lbl":
        nwave, ncol, nlyr, nstr = nwave or 100000, ncol or 1, nlyr or 40, 4
        wn = 500.0 + 0.1 * np.arange(nwave)
        centers = r.uniform(wn[0], wn[-1], 200)
        strength = 10.0 ** r.uniform(-3.0, 2.0, 200)
        width = r.uniform(0.05, 0.5, 200)
        kap = np.full(nwave, 1e-6)
        for c, s, w in zip(centers, strength, width):
            kap += s * w * w / ((wn - c) ** 2 + w * w)
        prof = np.exp(-4.0 * np.linspace(0.0, 1.0, nlyr))
        tau = np.clip(kap[:, None, None] * prof[None, None, :] * 10.0 ** r.uniform(-0.2, 0.2, (1, ncol, 1)),
                      1e-6, 1e2)
        ssa = r.uniform(0.0, 0.5, (nwave, ncol, nlyr))
        prop = np.stack([tau, ssa], axis=-1)
        bc = {"fbeam": np.ones((nwave, ncol)), "umu0": np.full((nwave, ncol), 0.6),
              "albedo": np.full((nwave, ncol), 0.3)}
        return Workload(name, nstr, "lamber,quiet,onlyfl", prop, bc)

    if name == "gcm":
        nwave, ncol, nlyr, nstr = nwave or 32, ncol or 65536, nlyr or 60, 16
        prop = np.zeros((nwave, ncol, nlyr, 2 + nstr))
        gas = 10.0 ** r.uniform(-5.0, 1.0, (nwave, 1, 1)) * 10.0 ** r.uniform(-0.5, 0.5, (nwave, ncol, nlyr))
        lo, hi = nlyr // 3, 2 * nlyr // 3
        aer = np.zeros((nwave, ncol, nlyr))
        aer[:, :, lo:hi] = 10.0 ** r.uniform(-2.0, 0.0, (1, ncol, 1))
        ssa_aer = r.uniform(0.9, 0.9999, (nwave, ncol, nlyr))
        tau = gas + aer
        prop[..., 0] = tau
        prop[..., 1] = np.clip(aer * ssa_aer / tau, 0.0, 0.9999)
        gg = r.uniform(0.6, 0.8, (1, ncol, 1))
        for k in range(nstr):
            prop[..., 2 + k] = np.where(aer > 0, gg ** (k + 1), 0.0)
        mu0 = r.uniform(0.05, 1.0, (1, ncol))
        # keep the beam away from the quadrature angles (DISORT rejects |mu0-mu_i|/mu0 < 1e-4)
        x, _ = np.polynomial.legendre.leggauss(nstr // 2)
        for q in 0.5 * (x + 1.0):
            mu0 = np.where(np.abs(mu0 - q) / mu0 < 1e-3, mu0 * 1.002, mu0)
        mu0 = np.minimum(mu0, 1.0)
        bc = {"fbeam": np.full((nwave, ncol), 1360.0 / nwave), "umu0": np.repeat(mu0, nwave, 0),
              "albedo": np.repeat(r.uniform(0.0, 0.4, (1, ncol)), nwave, 0)}
        return Workload(name, nstr, "lamber,quiet,onlyfl", prop, bc)

    raise KeyError(name)
