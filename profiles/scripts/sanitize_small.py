"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): flux solves at nstr 4/8/16 with
beam + thermal sources on the host and device paths, the general (intensity) kernel, band reduction and the
opacity kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pyharp_b200 import Disort, DisortOptions, opacity as OP, workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
for name, kw in (("lbl", dict(nwave=300, ncol=1, nlyr=11)), ("amars_sw", dict(nwave=37, ncol=3, nlyr=9)),
                 ("amars_lw", dict(nwave=9, ncol=5, nlyr=13)), ("gcm", dict(nwave=3, ncol=50, nlyr=7))):
    wl = W.make(name, **kw)
    ds = Disort(wl.options())
    t = lambda a, d: None if a is None else torch.from_numpy(a).to(d)
    for d in ("cpu", dev):
        flx = ds.forward(t(wl.prop, d), {k: t(v, d) for k, v in wl.bc.items()}, t(wl.temf, d))
        assert bool(torch.isfinite(flx).all())
    ds.reduce_band(flx, torch.ones(wl.shape[0], dtype=torch.float64))
    print(name, "ok", tuple(flx.shape))

import intensity_cases as IC  # noqa: E402
kw = IC.make((8, 5, 8, 1, 1, 1, 1, 1))
nw, nc, nl, _ = kw["prop"].shape
op = DisortOptions().flags(IC.flag_tokens(kw["flags"])).nwave(nw).ncol(nc)
op.ds().nlyr, op.ds().nstr, op.ds().nmom = nl, kw["nstr"], kw["nmom"]
op.user_phi(kw["phi"]).user_mu(kw["umu"]).user_tau(kw["utau"]).wave_lower(list(kw["wave_lower"])).wave_upper(list(kw["wave_upper"]))
ds = Disort(op)
tt = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ds.forward(tt(kw["prop"]), {k: tt(v) for k, v in kw["bc"].items()}, tt(kw["temf"]))
print("intensity ok", tuple(ds.get_rad().shape))

z = np.load(os.path.join(ROOT, "tests", "golden", "opacity_aerosol.npz"))
o = OP.AttenuatorOptions().species_weights(list(z["species_weights"])).species_ids([0])
a = OP.S8Fuller(o, table=z["s8_fuller_table"]).forward(torch.from_numpy(z["conc"]), {"wavenumber": torch.from_numpy(z["wavenumber"])})
OP.band_optics([a], torch.ones(a.shape[2], dtype=torch.float64))
OP.layer2level(torch.rand(4, 9, dtype=torch.float64, device=dev) + 1.0)
print("opacity ok")
torch.cuda.synchronize()

# round 2: the band-level call (fused optics, fused band integration with L2 reductions, warp-reduced form with
# replicas, chunked host entry, heating rate) and the solver-level fused band integration
from pyharp_b200 import RadiationBand, RadiationBandOptions  # noqa: E402
bw = W.make_band("amars_lw", nwave=12, ncol=37, nlyr=24)
band = W.band_of(bw)
band.options.host_chunk_cols(8)
band = RadiationBand(band.options, nspecies=2)
f = torch.from_numpy
rho = bw.pres / (188.9 * bw.temp)
for d in ("cpu", dev):
    mv = lambda a: f(a).to(d)
    out, dT = band.forward_band(mv(bw.conc), mv(bw.dz), {k: mv(v) for k, v in bw.bc.items()}, {"pres": mv(bw.pres), "temp": mv(bw.temp)},
                                rho=mv(rho), cp=844.0)
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(dT).all())
flx = band.forward(f(bw.conc).to(dev), f(bw.dz).to(dev), {k: f(v).to(dev) for k, v in bw.bc.items()},
                   {"pres": f(bw.pres).to(dev), "temp": f(bw.temp).to(dev)})
print("band ok", tuple(out.shape), tuple(flx.shape))
for name, kw in (("lbl", dict(nwave=500, ncol=1, nlyr=9)), ("lbl", dict(nwave=200, ncol=4, nlyr=9)), ("amars_lw", dict(nwave=8, ncol=40, nlyr=9)),
                 ("gcm", dict(nwave=3, ncol=40, nlyr=5))):
    wl = W.make(name, **kw)
    ds = Disort(wl.options())
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    b = ds.forward_band(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf), torch.ones(wl.shape[0], dtype=torch.float64))
    assert bool(torch.isfinite(b).all())
    print("solve_band", name, kw["ncol"], "ok")
torch.cuda.synchronize()
