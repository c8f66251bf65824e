#!/usr/bin/env python
"""Where a small call's ~0.1 ms goes: host-side issue rate (calls back to back, one synchronise at the end) against the
synchronised per-call latency, and the device time of the same calls from CUDA events."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from pyharp_b200 import Disort, workloads as W

dev = torch.device("cuda", 0)
for name, kw in (("amars_lw", dict(nwave=32, ncol=1, nlyr=40)), ("lbl", dict(nwave=12500, ncol=1, nlyr=40))):
    wl = W.make(name, **kw)
    t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
    dp, dbc, dt_ = t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)
    w = torch.ones(wl.shape[0], dtype=torch.float64, device=dev)
    ds = Disort(wl.options())
    for fn_name, fn in (("forward", lambda: ds.forward(dp, dbc, dt_)), ("forward_band(check=False)", lambda: ds.forward_band(dp, dbc, dt_, w, check=False))):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        n = 300
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
            torch.cuda.synchronize()
        sync_ms = (time.perf_counter() - t0) / n * 1e3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        issue_ms = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        print(json.dumps({"case": f"{name} {wl.nsolve} solves", "call": fn_name, "synchronised_ms": round(sync_ms, 4),
                          "host_issue_ms": round(issue_ms, 4), "device_ms_back_to_back": round(e0.elapsed_time(e1) / n, 4)}), flush=True)
