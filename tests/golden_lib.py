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


def diverr(a, ref, mask=None):
    """Flux-DIVERGENCE criterion: |(net_a[k+1] - net_a[k]) - (net_ref[k+1] - net_ref[k])| per layer, relative to the
    largest flux of the solve, net = up - down.  This is the quantity pyharp consumes -- the heating rate is the
    difference of the net flux across a layer (examples/amars_sw.cpp:290-301) -- and in an optically thick column
    it is orders of magnitude below the fluxes themselves, so the plain flux criterion does not bound it."""
    na, nr = a[..., 0] - a[..., 1], ref[..., 0] - ref[..., 1]
    e = np.abs(np.diff(na, axis=2) - np.diff(nr, axis=2)) / np.max(np.abs(ref), axis=(2, 3))[..., None]
    if mask is not None:
        e = e[mask]
    return float(np.max(e)) if e.size else 0.0


def elemerr(a, ref, floor=1e-6):
    """Largest ELEMENTWISE relative error, with an absolute floor of `floor` x the largest flux of the solve in the
    denominator (fluxes below that are compared absolutely)."""
    top = np.max(np.abs(ref), axis=(2, 3), keepdims=True)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor * top)))


def check(a, ref, rtol=1e-9, mask=None):
    """The parity bar of this project on one batch of fluxes: flux error and flux-divergence error both within
    rtol of the largest flux of the solve; returns (flux error, divergence error, elementwise error) for reports."""
    fe, de = relerr(a, ref, mask), diverr(a, ref, mask)
    assert fe < rtol, f"flux error {fe:.3e} >= {rtol}"
    assert de < rtol, f"flux-divergence error {de:.3e} >= {rtol}"
    return fe, de, elemerr(a, ref)


def disotest_cases():
    """tests/golden/disotest.json (make_golden_disotest.py) as (name, Workload, flux (1, 1, nlev, 2), case dict)."""
    import json
    out = []
    for c in json.load(open(os.path.join(GOLDEN_DIR, "disotest.json"))):
        L, N = len(c["dtau"]), c["nstr"]
        pm = np.zeros(N)
        if c["phase"] == "rayleigh":
            pm[1] = 0.1
        elif c["phase"] == "hg":
            pm[:] = c["g"] ** np.arange(1, N + 1)
        prop = np.zeros((1, 1, L, 2 + N))
        prop[0, 0, :, 0], prop[0, 0, :, 1], prop[0, 0, :, 2:] = c["dtau"][::-1], c["ssa"][::-1], pm[None, :]
        one = lambda v: np.full((1, 1), float(v))
        bc = {"albedo": one(c["albedo"])}
        tokens, temf, wl_, wu_ = "lamber,quiet,onlyfl", None, None, None
        if c["fbeam"]:
            bc.update(umu0=one(c["umu0"]), fbeam=one(c["fbeam"]))
        if c["fisot"]:
            bc["fisot"] = one(c["fisot"])
        if c["temper"] is not None:
            tokens += ",planck"
            temf = np.array(c["temper"][::-1])[None, :]
            bc.update(btemp=one(c["btemp"]), ttemp=one(c["ttemp"]), temis=one(c["temis"]))
            wl_, wu_ = np.array([c["wvn"][0]]), np.array([c["wvn"][1]])
        out.append((c["name"], W.Workload("disotest_" + c["name"], N, tokens, prop, bc, temf, wl_, wu_),
                    np.array(c["flux"])[None, None], c))
    return out
