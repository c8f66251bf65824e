#!/usr/bin/env python
"""e45: per-call latency of small nstr = 8 batches on the three forms of the flux solve (fused thread-per-solve kernel,
warp-per-solve form, layer-split form with the team adding kernel), device tensors, synchronised after every call."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from pyharp_b200 import Disort, workloads as W


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dev = torch.device("cuda", 0)
import os
CASES = [("amars_sw", 40, (32, 128, 500, 1184, 2368, 4736, 9472, 18944, 37888)), ("amars_lw", 80, (32, 128, 500, 1184, 2368, 4736, 9472, 18944, 37888)),
         ("lbl", 40, (1184, 4736, 12500, 25000, 50000, 100000))]
if os.environ.get("CASES"):
    CASES = [c for c in CASES if c[0] in os.environ["CASES"].split(",")]
for name, nlyr, sizes in CASES:
    for nwave in sizes:
        wl = W.make(name, nwave=nwave, ncol=1, nlyr=nlyr)
        t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
        dp, dbc, dt_ = t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)
        row = {"case": f"{name} {nwave} x 1 x {nlyr}"}
        for form, (warp, split) in (("fused", (0, 0)), ("warp", (1 << 30, 0)), ("split", (0, 1))):
            ds = Disort(wl.options()).set_warp_solve_max(warp).set_layer_split(split)
            row[form + "_ms"] = round(timeit(lambda: ds.forward(dp, dbc, dt_), 30), 4)
            if name == "lbl":  # the spectral-point sharding's step: solve + fused band integration (one column: warp-reduced form)
                w = torch.ones(nwave, dtype=torch.float64, device=dev)
                row[form + "_band_ms"] = round(timeit(lambda: ds.forward_band(dp, dbc, dt_, w, check=False), 30), 4)
        print(json.dumps(row), flush=True)
