"""Layer-split form (nstr = 16: prefix / layer / team adding kernels; nstr = 24: thread form of the adding phase) on a batch
that does not fill its last group of 8 solves, for compute-sanitizer (memcheck, racecheck, synccheck)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyharp_b200 import Disort, workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
wl = W.make("gcm", nwave=3, ncol=37, nlyr=9)
t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
res = []
for split in (0, 1):
    ds = Disort(wl.options()).set_layer_split(split)
    res.append(ds.forward(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)).cpu().numpy())
    w = torch.ones(wl.shape[0], dtype=torch.float64, device=dev)
    ds.forward_band(t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf), w, check=True)
print("max |split - fused| / max|F| =", float(np.max(np.abs(res[0] - res[1])) / np.max(np.abs(res[0]))))
