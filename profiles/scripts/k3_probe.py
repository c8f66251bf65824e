#!/usr/bin/env python
"""tests/test_disort.cpp's own case (10 x 1 solves, one layer, nstr 16, intensities at 6 angles x 2 levels): per-call time of
the general kernel K3 next to the single-thread oracle, and where a call's time goes (kernel vs host side)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle_lib as O
from pyharp_b200 import Disort, DisortOptions

dev = torch.device("cuda", 0)
op = DisortOptions().flags("usrtau,usrang,lamber,quiet,intensity_correction,old_intensity_correction").nwave(10).ncol(1)
op.user_mu([-1, -0.5, -0.1, 0.1, 0.5, 1]).user_phi([0]).user_tau([0, 0.03125])
op.ds().nlyr, op.ds().nstr, op.ds().nmom = 1, 16, 16
ds = Disort(op)
prop = torch.zeros((10, 1, 1, 18), dtype=torch.float64)
prop[..., 0], prop[..., 1] = 0.03125, 0.2
bc = {"umu0": torch.full((10, 1), 0.1, dtype=torch.float64), "fbeam": torch.full((10, 1), np.pi / 0.1, dtype=torch.float64)}
dprop, dbc = prop.to(dev), {k: v.to(dev) for k, v in bc.items()}


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


host = timeit(lambda: ds.forward(prop, bc))
devt = timeit(lambda: ds.forward(dprop, dbc))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ds.forward(dprop, dbc)
e1.record()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    O.forward(prop.numpy(), {k: v.numpy() for k, v in bc.items()}, None, nstr=16, flags=O.USRTAU | O.USRANG | O.LAMBER | O.INTCOR | O.OLDINTCOR,
              umu=np.array([-1, -0.5, -0.1, 0.1, 0.5, 1.0]), phi=np.array([0.0]), utau=np.array([0, 0.03125]), nthreads=1)
cpu = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps({"case": "test_disort 10 x 1 x 1, nstr 16, intensities", "forward_host_ms": round(host, 4), "forward_device_ms": round(devt, 4),
                  "device_time_per_call_ms": round(e0.elapsed_time(e1) / 20, 4), "oracle_1thread_ms": round(cpu, 3)}))
