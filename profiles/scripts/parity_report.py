#!/usr/bin/env python
"""Parity report on the GPU: for every committed fixture, every DISOTEST configuration and a seeded sample of each
BASELINE.json workload shape, the CUDA path (through the C ABI) against the 80-bit oracle:
  flux      max |F - F_ref| / max|F| per solve             (bar 1e-9, BASELINE.json north_star)
  diverg    max |d(net) - d(net_ref)| / max|F| per layer   (bar 1e-9; what the heating rate consumes)
  elem      max elementwise relative error with an absolute floor of 1e-6 max|F|
and the double-precision oracle's own distance to its 80-bit build beside it.
  python profiles/scripts/parity_report.py > profiles/r02/parity_report.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import golden_lib as G
import oracle_lib as O
from pyharp_b200 import Disort, workloads as W


def gpu(wl):
    ds = Disort(wl.options())
    t = lambda a: None if a is None else torch.from_numpy(a).cuda()
    return ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)).cpu().numpy()


def orc(wl, ld):
    f, _, st = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags, wave_lower=wl.wave_lower,
                         wave_upper=wl.wave_upper, nthreads=os.cpu_count() or 8, long_double=ld)
    assert not st.any()
    return f


def row(name, wl, ref):
    a = gpu(wl)
    d = orc(wl, False)
    print(f"{name:34s} {wl.nsolve:8d} {wl.nstr:4d} {wl.shape[2]:4d}   {G.relerr(a, ref):9.2e} {G.diverr(a, ref):9.2e} {G.elemerr(a, ref):9.2e}"
          f"   | {G.relerr(d, ref):9.2e} {G.diverr(d, ref):9.2e}")


def main():
    torch.cuda.set_device(0)
    print("CUDA path vs the 80-bit oracle (and the FP64 oracle vs the same)   --   " + torch.cuda.get_device_name(0))
    print(f"{'case':34s} {'solves':>8s} {'nstr':>4s} {'nlyr':>4s}   {'flux':>9s} {'diverg':>9s} {'elem':>9s}   | {'fp64 flux':>9s} {'fp64 div':>9s}")
    for name in G.NAMES:
        wl, ld, _, _ = G.load(name)
        row("golden/" + name, wl, ld)
    for name, wl, want, _ in G.disotest_cases():
        row("disotest/" + name, wl, want)
    for name, kw in (("amars_sw", dict(nwave=500, ncol=8)), ("amars_lw", dict(nwave=32, ncol=64)),
                     ("amars_lw", dict(nwave=32, ncol=64, variant=1)), ("lbl", dict(nwave=20000, ncol=1)),
                     ("gcm", dict(nwave=8, ncol=128))):
        wl = W.make(name, **kw)
        row(f"workload/{name}{'/v1' if kw.get('variant') else ''}", wl, orc(wl, True))
    # the band-level workload (realistic pressure / temperature profile: thin upper layers)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import band_chain as BC
    bw = W.make_band("amars_lw", nwave=32, ncol=64, nlyr=80)
    prop, temf, flx_ref, bflx_ref = BC.lw_chain(bw)
    wl = W.Workload("band_lw", 8, bw.flag_tokens, prop, bw.bc_full(), temf, bw.wave_lower, bw.wave_upper)
    row("workload/amars_lw band (RFM optics)", wl, flx_ref)
    band = W.band_of(bw)
    t = lambda a: torch.from_numpy(a).cuda()
    b = band.forward_band(t(bw.conc), t(bw.dz), {k: t(v) for k, v in bw.bc.items()}, {"pres": t(bw.pres), "temp": t(bw.temp)}).cpu().numpy()
    net, nr = b[..., 0] - b[..., 1], bflx_ref[..., 0] - bflx_ref[..., 1]
    print(f"{'band flux (fused integration)':34s} {bw.nsolve:8d} {8:4d} {80:4d}   {np.max(np.abs(b - bflx_ref)) / np.max(np.abs(bflx_ref)):9.2e} "
          f"{np.max(np.abs(np.diff(net, axis=1) - np.diff(nr, axis=1))) / np.max(np.abs(bflx_ref)):9.2e}")


if __name__ == "__main__":
    main()
