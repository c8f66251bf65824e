"""Loads tests/golden/*.npz fixtures (written by tests/golden/make_golden.py) as Workloads."""
import glob
import os

import numpy as np

from pyharp_b200 import workloads as W

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# solver fixtures (make_golden.py, make_golden_amars.py); opacity_*.npz belong to make_golden_opacity.py
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
               if not os.path.basename(p).startswith("opacity_"))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    bc = {k[3:]: z[k] for k in z.files if k.startswith("bc_")}
    wl = W.Workload(name, int(z["nstr"]), str(z["flag_tokens"]), z["prop"], bc,
                    z["temf"] if "temf" in z.files else None,
                    z["wave_lower"] if "wave_lower" in z.files else None,
                    z["wave_upper"] if "wave_upper" in z.files else None)
    ind = z["flx_independent"] if "flx_independent" in z.files else None
    return wl, z["flx_oracle_ld"], ind, z["independent_valid"]


def relerr(a, ref, mask=None):
    """max |a - ref| relative to the largest flux of each solve (the 1e-9 criterion of the north star)"""
    e = np.abs(a - ref) / np.max(np.abs(ref), axis=(2, 3), keepdims=True)
    if mask is not None:
        e = e[mask]
    return float(np.max(e))
