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


# ---------------------------------------------------------------------------------------------
# Band-level workloads: the inputs of RadiationBand.forward (conc / pres / temp + tables) rather than
# the optical properties, so that the opacity kernels are on the measured path too.
# ---------------------------------------------------------------------------------------------
@dataclass
class BandWorkload:
    name: str
    nstr: int
    flag_tokens: str
    nwave: int
    tables: list            # per attenuator: dict(kind="rfm", species_id, kaxis, kdata, krefatm)
    conc: np.ndarray        # (ncol, nlyr, nspecies) mol/m^3, layer 0 = surface
    pres: np.ndarray        # (ncol, nlyr) Pa
    temp: np.ndarray        # (ncol, nlyr) K
    dz: np.ndarray          # (nlyr) m
    bc: dict                # name -> (ncol): shared by all spectral points of the band
    weights: np.ndarray     # (nwave)
    wave_lower: np.ndarray = None
    wave_upper: np.ndarray = None

    @property
    def ncol(self):
        return self.conc.shape[0]

    @property
    def nlyr(self):
        return self.conc.shape[1]

    @property
    def nsolve(self):
        return self.nwave * self.ncol

    def bc_full(self):
        """The boundary fields in pydisort's (nwave, ncol) shape."""
        return {k: np.ascontiguousarray(np.broadcast_to(v[None, :], (self.nwave, self.ncol))) for k, v in self.bc.items()}

    def h2d_bytes(self):
        return self.conc.nbytes + self.pres.nbytes + self.temp.nbytes + self.dz.nbytes + sum(v.nbytes for v in self.bc.values())


def make_band(name="amars_lw", nwave=None, ncol=None, nlyr=None, variant=0):
    """amars_lw as a band (examples/amars_lw.cpp:40-88 with a real atmosphere instead of its constants): two RFM
    correlated-k attenuators (CO2, H2O) on one synthetic table file (`amarsw-ck-B1.nc` is absent from the reference
    tree), nstr = 8, thermal emission in [1, 150] cm^-1, Lambertian surface.  The table has npres >= nlyr rows because
    RFM's get_reftemp reads the first nlyr entries of the reference atmosphere (src/opacity/rfm.cpp:178)."""
    if name != "amars_lw":
        raise KeyError(name)
    nwave, ncol, nlyr = nwave or 32, ncol or 1, nlyr or 80
    r = _rng("amars_lw", 50 + variant)
    npres, ntemp = max(96, nlyr), 8
    R = 8.314462618
    # table axes: g-points, ln p descending from the surface up (like the layers), temperature anomaly
    g = (np.arange(nwave) + 0.5) / nwave
    lnp_ax = np.linspace(np.log(1.2e5), np.log(5.0), npres)
    dT_ax = np.linspace(-70.0, 70.0, ntemp)
    tref = np.linspace(290.0, 150.0, npres) + 5.0 * np.sin(np.linspace(0.0, 6.0, npres))
    kaxis = np.concatenate([g, lnp_ax, dT_ax])
    krefatm = np.stack([lnp_ax, tref])
    tables = []
    for sid, (lo, hi, pexp, tcoef) in enumerate(((-13.0, -1.5, 0.45, 0.012), (-9.0, 4.0, 0.3, 0.02))):
        base = lo + (hi - lo) * g ** 1.5
        kd = (base[:, None, None] + pexp * (lnp_ax[None, :, None] - np.log(1.0e5)) + tcoef * dT_ax[None, None, :]
              + 0.05 * r.standard_normal((nwave, npres, ntemp)))
        tables.append(dict(kind="rfm", species_id=sid, kaxis=kaxis, kdata=np.ascontiguousarray(kd), krefatm=krefatm))
    # atmosphere: hydrostatic-like pressure, lapse-rate temperature with a per-column offset (index 0 = surface)
    z = (np.arange(nlyr) + 0.5) / nlyr
    p0 = 1.0e5 * (1.0 + 0.05 * r.uniform(-1.0, 1.0, (ncol, 1)))
    pres = p0 * np.exp(-7.5 * z[None, :])
    temp = 295.0 - 140.0 * z[None, :] + r.uniform(-6.0, 6.0, (ncol, 1)) + 1.5 * r.standard_normal((ncol, nlyr))
    ntot = pres / (R * temp)
    q = 10.0 ** r.uniform(-4.0, -2.5, (ncol, 1)) * np.exp(-3.0 * z[None, :])
    conc = np.stack([0.95 * ntot, q * ntot], axis=-1)
    dz = np.full(nlyr, 60.0e3 / nlyr)
    tsurf = temp[:, 0] + 2.0
    bc = {"albedo": r.uniform(0.0, 0.3, ncol) if variant else np.zeros(ncol), "btemp": tsurf}
    w = r.uniform(0.5, 1.5, nwave)
    return BandWorkload(name, 8, "lamber,quiet,onlyfl,planck", nwave, tables, np.ascontiguousarray(conc),
                        np.ascontiguousarray(pres), np.ascontiguousarray(temp), dz, bc, w / w.sum(),
                        np.full(nwave, 1.0), np.full(nwave, 150.0))


def band_of(bw, ncol=None, exact_interp=False, mix_eps=0.0, device=None):
    """pyharp_b200.RadiationBand for a BandWorkload."""
    from .band import RadiationBand, RadiationBandOptions
    from .disort import DisortOptions
    from .opacity import AttenuatorOptions, RFM
    atts = {}
    for k, t in enumerate(bw.tables):
        op = AttenuatorOptions().species_ids([t["species_id"]]).species_names(["CO2", "H2O"]).type("rfm")
        atts[f"a{k}"] = RFM(op, t["kaxis"], t["kdata"], t["krefatm"], device=device or "cuda")
    dop = DisortOptions().flags(bw.flag_tokens)
    d = dop.ds()
    d.nstr = d.nmom = bw.nstr
    bo = (RadiationBandOptions().attenuators(atts).disort(dop).nlyr(bw.nlyr).ncol(ncol or bw.ncol)
          .wave_lower(list(bw.wave_lower)).wave_upper(list(bw.wave_upper)).weights(bw.weights)
          .exact_interp(exact_interp).mix_eps(mix_eps))
    return RadiationBand(bo, device=device, nspecies=bw.conc.shape[2])
