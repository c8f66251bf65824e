#!/usr/bin/env python
"""Per-call latency of the small batches pyharp's own drivers issue (examples/amars_sw.cpp:200-208: 500 x 1;
examples/amars_lw.cpp:51-54: G x 1; tests/test_disort.cpp: 10 x 1 with intensities), synchronised after every call,
next to the single-thread oracle (the reference's execution model: a serial loop over c_disort).
  python profiles/scripts/latency_probe.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib as O
from pyharp_b200 import Disort, DisortOptions, workloads as W


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    cases = [("amars_sw 500 x 1 x 40, nstr 8", W.make("amars_sw", nwave=500, ncol=1, nlyr=40)),
             ("amars_lw 32 x 1 x 80, nstr 8", W.make("amars_lw", nwave=32, ncol=1, nlyr=80)),
             ("amars_lw 32 x 1 x 40, nstr 8", W.make("amars_lw", nwave=32, ncol=1, nlyr=40)),
             ("lbl 12500 x 1 x 40, nstr 4 (1/8 of config 4)", W.make("lbl", nwave=12500, ncol=1, nlyr=40)),
             ("lbl 100000 x 1 x 40, nstr 4", W.make("lbl", nwave=100000, ncol=1, nlyr=40)),
             ("gcm 32 x 1 x 60, nstr 16", W.make("gcm", nwave=32, ncol=1, nlyr=60))]
    for name, wl in cases:
        ds = Disort(wl.options())
        t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
        dp, dbc, dt_ = t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)
        w = torch.ones(wl.shape[0], dtype=torch.float64, device=dev)
        lat_dev = timeit(lambda: ds.forward(dp, dbc, dt_), 50)
        lat_band = timeit(lambda: ds.forward_band(dp, dbc, dt_, w, check=False), 50)
        hp, hbc, ht = torch.from_numpy(wl.prop), {k: torch.from_numpy(v) for k, v in wl.bc.items()}, None if wl.temf is None else torch.from_numpy(wl.temf)
        lat_host = timeit(lambda: ds.forward(hp, hbc, ht), 30)
        st = ds.stats()
        nsample = min(wl.nsolve, 512)
        sub = W.Workload(wl.name, wl.nstr, wl.flag_tokens, wl.prop[:nsample], {k: v[:nsample] for k, v in wl.bc.items()}, wl.temf,
                         None if wl.wave_lower is None else wl.wave_lower[:nsample], None if wl.wave_upper is None else wl.wave_upper[:nsample])
        t0 = time.perf_counter()
        O.forward(sub.prop, sub.bc, sub.temf, nstr=sub.nstr, flags=sub.flags, wave_lower=sub.wave_lower, wave_upper=sub.wave_upper, nthreads=1)
        cpu_ms = (time.perf_counter() - t0) * 1e3 * wl.nsolve / nsample
        print(json.dumps({"case": name, "solves": wl.nsolve, "forward_device_ms": round(lat_dev, 4), "solve_band_device_ms": round(lat_band, 4),
                          "forward_host_ms": round(lat_host, 4), "oracle_1thread_ms": round(cpu_ms, 2),
                          "speedup_host_call_vs_1thread": round(cpu_ms / lat_host, 1)}))
    # test_disort.cpp: 10 x 1, one layer, nstr 16, intensities at 6 angles x 2 levels (general kernel K3)
    op = DisortOptions().flags("usrtau,usrang,lamber,quiet,intensity_correction,old_intensity_correction").nwave(10).ncol(1)
    op.user_mu([-1, -0.5, -0.1, 0.1, 0.5, 1]).user_phi([0]).user_tau([0, 0.03125])
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = 1, 16, 16
    ds = Disort(op)
    prop = torch.zeros((10, 1, 1, 18), dtype=torch.float64)
    prop[..., 0], prop[..., 1] = 0.03125, 0.2
    bc = {"umu0": torch.full((10, 1), 0.1, dtype=torch.float64), "fbeam": torch.full((10, 1), np.pi / 0.1, dtype=torch.float64)}
    lat = timeit(lambda: ds.forward(prop, bc), 30)
    print(json.dumps({"case": "test_disort 10 x 1 x 1, nstr 16, intensities (K3)", "solves": 10, "forward_host_ms": round(lat, 4)}))


if __name__ == "__main__":
    main()
