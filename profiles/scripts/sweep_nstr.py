"""Throughput of the flux kernel for every compiled stream count on one workload family: 60 layers, beam source,
Henyey-Greenstein aerosol in the middle third of the column (all nstr phase moments supplied), gas absorption
elsewhere -- the GCM workload of bench.py with nstr varied.  Kernel time from CUDA events (Disort.stats())."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyharp_b200 import Disort, DisortOptions  # noqa: E402

dev = torch.device("cuda", 0)
nlyr, nwave = 60, 8
for nstr in [int(x) for x in os.environ.get("NSTRS", "2,4,6,8,12,16,24,32").split(",")]:
    ncol = {2: 65536, 4: 65536, 6: 32768, 8: 32768, 12: 16384, 16: 8192, 24: 2048, 32: 1024}[nstr]
    r = np.random.default_rng(nstr)
    prop = np.zeros((nwave, ncol, nlyr, 2 + nstr))
    gas = 10.0 ** r.uniform(-5.0, 1.0, (nwave, 1, 1)) * 10.0 ** r.uniform(-0.5, 0.5, (nwave, ncol, nlyr))
    aer = np.zeros((nwave, ncol, nlyr))
    aer[:, :, nlyr // 3:2 * nlyr // 3] = 10.0 ** r.uniform(-2.0, 0.0, (1, ncol, 1))
    tau = gas + aer
    prop[..., 0] = tau
    prop[..., 1] = np.clip(aer * r.uniform(0.9, 0.9999, tau.shape) / tau, 0.0, 0.9999)
    g = r.uniform(0.6, 0.8, (1, ncol, 1))
    for k in range(nstr):
        prop[..., 2 + k] = np.where(aer > 0, g ** (k + 1), 0.0)
    mu0 = r.uniform(0.05, 1.0, (nwave, ncol))
    x, _ = np.polynomial.legendre.leggauss(nstr // 2)
    for q in 0.5 * (x + 1.0):
        mu0 = np.where(np.abs(mu0 - q) / mu0 < 1e-3, mu0 * 1.003, mu0)
    bc = {"umu0": np.minimum(mu0, 1.0), "fbeam": np.ones((nwave, ncol)), "albedo": np.full((nwave, ncol), 0.3)}
    op = DisortOptions().flags("lamber,quiet,onlyfl").nwave(nwave).ncol(ncol)
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = nlyr, nstr, nstr
    ds = Disort(op)
    t = lambda a: torch.from_numpy(a).to(dev)
    dp, dbc = t(prop), {k: t(v) for k, v in bc.items()}
    for _ in range(2):
        flx = ds.forward(dp, dbc)
    torch.cuda.synchronize()
    assert int(ds.status().abs().sum()) == 0 and bool(torch.isfinite(flx).all())
    ds.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 4 if nstr <= 16 else 1
    e0.record()
    for _ in range(reps):
        ds.forward(dp, dbc)
    e1.record()
    torch.cuda.synchronize()
    st = ds.stats()
    call_ms = e0.elapsed_time(e1) / reps   # whole forward() call on device tensors (general kernel: not in stats)
    ms = st["solve_kernel_ms"] / max(1, st["solves_timed"]) if st["solves_timed"] else call_ms
    print(json.dumps({"nstr": nstr, "nlyr": nlyr, "solves": nwave * ncol, "kernel_ms": round(ms, 3), "call_ms": round(call_ms, 3),
                      "solves_per_s": round(nwave * ncol / (ms * 1e-3), 0),
                      "general_kernel": bool(os.environ.get("B200DISORT_FORCE_GENERAL"))}), flush=True)
    del ds, dp, dbc, flx
    torch.cuda.empty_cache()
