#!/usr/bin/env python
"""examples/amars_lw.cpp of pyharp as ONE band-level call on the B200.

The reference example builds two RFM correlated-k attenuators (CO2, H2O), sums their optical properties, calls
layer2level, runs disort->forward and integrates the fluxes with the ck weights -- five steps on the host
(examples/amars_lw.cpp:40-88).  Here the same pipeline is `RadiationBand.forward_band`: conc / pres / temp in, band
flux (ncol, nlyr+1, 2) and heating rate out; the optical properties and the per-g-point fluxes never leave the GPU.
The table file `amarsw-ck-B1.nc` is not part of the reference tree, so a synthetic ck table stands in for it.

    python examples/amars_lw_band.py [ncol]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pyharp_b200 import DisortOptions, RadiationBand, RadiationBandOptions, opacity as OP, workloads as W


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    bw = W.make_band("amars_lw", nwave=32, ncol=ncol, nlyr=80)        # synthetic table + atmosphere (seeded)
    # attenuators: harp::RFM (src/opacity/rfm.hpp) -- kaxis = [wavenumber | ln p | dT], kdata = ln sigma, krefatm
    atts = {}
    for name, t in zip(("CO2", "H2O"), bw.tables):
        op = OP.AttenuatorOptions().type("rfm").species_names(["CO2", "H2O"]).species_ids([t["species_id"]])
        atts[name] = OP.RFM(op, t["kaxis"], t["kdata"], t["krefatm"])
    # solver options exactly as disort_options_lw() sets them (examples/amars_lw.cpp:18-38)
    dop = DisortOptions().header("running amars lw").flags("lamber,quiet,onlyfl,planck,intensity_correction,old_intensity_correction")
    dop.ds().nstr = dop.ds().nmom = 8
    bo = (RadiationBandOptions().name("B1").attenuators(atts).disort(dop).nlyr(bw.nlyr).ncol(ncol)
          .wave_lower([1.0] * bw.nwave).wave_upper([150.0] * bw.nwave).weights(bw.weights))
    band = RadiationBand(bo, nspecies=2)

    f = torch.from_numpy
    conc, dz = f(bw.conc), f(bw.dz)                                     # CPU tensors, as the reference's callers pass
    bc = {"albedo": f(bw.bc["albedo"]), "btemp": f(bw.bc["btemp"])}     # per column: shared by the band's g-points
    op = {"pres": f(bw.pres), "temp": f(bw.temp)}
    rho = f(bw.pres / (188.9 * bw.temp))
    bflx, dTdt = band.forward_band(conc, dz, bc, op, rho=rho, cp=844.0)
    t0 = time.perf_counter()
    for _ in range(5):
        bflx, dTdt = band.forward_band(conc, dz, bc, op, rho=rho, cp=844.0)
    dt = (time.perf_counter() - t0) / 5
    print(f"{bw.nwave} g-points x {ncol} columns x {bw.nlyr} layers: {dt * 1e3:.2f} ms per call "
          f"({bw.nwave * ncol / dt / 1e6:.1f} M DISORT solves/s, host tensors in and out)")
    print("column 0: band flux at the surface (up, down) =", bflx[0, 0].tolist(), " at the top =", bflx[0, -1].tolist())
    print("column 0: heating rate [K/day] in the lowest / highest layer =", float(dTdt[0, 0]) * 86400.0, float(dTdt[0, -1]) * 86400.0)
    # the reference's return value (per-g-point fluxes) and the optics it exports are available too
    flx = band.forward(conc, dz, {k: v[None, :].expand(bw.nwave, -1).contiguous() for k, v in bc.items()}, op)
    print("per-g-point flux", tuple(flx.shape), "; optics", tuple(band.shared["radiation/B1/optics"].shape))


if __name__ == "__main__":
    main()
