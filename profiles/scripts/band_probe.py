#!/usr/bin/env python
"""Timing probe of the band-level forward (device and host entry) next to the prop-level step; prints JSON lines.
  python profiles/scripts/band_probe.py [--ncol N] [--steps K] [--chunk C]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyharp_b200 import workloads as W


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncol", type=int, default=16384)
    ap.add_argument("--nwave", type=int, default=32)
    ap.add_argument("--nlyr", type=int, default=80)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--exact", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    bw = W.make_band("amars_lw", nwave=args.nwave, ncol=args.ncol, nlyr=args.nlyr)
    band = W.band_of(bw, exact_interp=args.exact)
    if args.chunk:
        band.options.host_chunk_cols(args.chunk)
        from pyharp_b200 import RadiationBand
        band = RadiationBand(band.options, nspecies=2)
    t = lambda a: torch.from_numpy(a).to(dev)
    d = dict(conc=t(bw.conc), dz=t(bw.dz), bc={k: t(v) for k, v in bw.bc.items()}, op={"pres": t(bw.pres), "temp": t(bw.temp)})
    for _ in range(3):
        out = band.forward_band(d["conc"], d["dz"], d["bc"], d["op"])
    torch.cuda.synchronize()
    band.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        band.forward_band(d["conc"], d["dz"], d["bc"], d["op"], check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    st = band.stats()
    kms = st["solve_kernel_ms"] / max(1, st["solves_timed"])
    print(json.dumps({"leg": "device", "ms_per_step": ms, "solves_per_s": bw.nsolve / ms * 1e3, "k1_ms": kms,
                      "k1_solves_per_s": bw.nsolve / kms * 1e3, "launches": st["kernel_launches"]}))
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h = dict(conc=pin(bw.conc), dz=pin(bw.dz), bc={k: pin(v) for k, v in bw.bc.items()}, op={"pres": pin(bw.pres), "temp": pin(bw.temp)})
    for _ in range(3):
        outh = band.forward_band(h["conc"], h["dz"], h["bc"], h["op"])
    assert float((outh - out.cpu()).abs().max() / out.abs().max()) < 1e-12
    t0 = time.perf_counter()
    for _ in range(args.steps):
        band.forward_band(h["conc"], h["dz"], h["bc"], h["op"])
    dt = (time.perf_counter() - t0) / args.steps
    st = band.stats()
    print(json.dumps({"leg": "host", "ms_per_step": dt * 1e3, "solves_per_s": bw.nsolve / dt, "h2d_bytes": bw.h2d_bytes(),
                      "d2h_bytes": outh.nbytes, "k1_launches_timed": st["solves_timed"], "k1_ms_sum_per_step": st["solve_kernel_ms"] / args.steps}))
    # pageable host memory
    hp = dict(conc=torch.from_numpy(bw.conc), dz=torch.from_numpy(bw.dz), bc={k: torch.from_numpy(v) for k, v in bw.bc.items()},
              op={"pres": torch.from_numpy(bw.pres), "temp": torch.from_numpy(bw.temp)})
    band.forward_band(hp["conc"], hp["dz"], hp["bc"], hp["op"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        band.forward_band(hp["conc"], hp["dz"], hp["bc"], hp["op"])
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"leg": "host_pageable_inputs", "ms_per_step": dt * 1e3, "solves_per_s": bw.nsolve / dt}))


if __name__ == "__main__":
    main()
