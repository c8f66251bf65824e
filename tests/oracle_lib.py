"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- test infrastructure only.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (pyharp_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")

USRTAU, USRANG, LAMBER, PLANCK, ONLYFL, INTCOR, OLDINTCOR = 1, 2, 4, 8, 16, 32, 64


class OrcBC(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in
                ("umu0", "fbeam", "albedo", "btemp", "ttemp", "temis", "fisot", "phi0")]


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _load(name):
    path = os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    lib.orc_forward.restype = C.c_int
    lib.orc_forward.argtypes = [C.c_int] * 7 + [C.c_double] + [C.c_int] * 3 + [dp] * 6 + \
        [C.POINTER(OrcBC), dp, dp, dp, C.POINTER(C.c_int), C.c_int]
    lib.orc_plkavg.restype = C.c_double
    lib.orc_plkavg.argtypes = [C.c_double] * 3
    lib.orc_qgausn.argtypes = [C.c_int, dp, dp]
    lib.orc_asymtx.restype = C.c_int
    lib.orc_asymtx.argtypes = [dp, C.c_int, dp, dp]
    lib.orc_lepoly.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp]
    return lib


_libs = {}


def lib(long_double=False):
    key = "liboracle_ld.so" if long_double else "liboracle.so"
    if key not in _libs:
        _libs[key] = _load(key)
    return _libs[key]


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _arr(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def forward(prop, bc=None, temf=None, *, nstr, nmom=None, flags=LAMBER | ONLYFL, accur=0.0,
            utau=None, umu=None, phi=None, wave_lower=None, wave_upper=None, nthreads=1,
            long_double=False):
    """pydisort-style forward on the CPU oracle.

    prop (nwave,ncol,nlyr,nprop), index 0 = surface.  Returns (flx, uu, status).
    """
    prop = _arr(prop)
    nwave, ncol, nlyr, nprop = prop.shape
    nmom = nstr if nmom is None else nmom
    utau, umu, phi = _arr(utau), _arr(umu), _arr(phi)
    ntau = 0 if utau is None else len(utau)
    numu = 0 if umu is None else len(umu)
    nphi = 0 if phi is None else len(phi)
    onlyfl = bool(flags & ONLYFL)
    ntau_eff = ntau if flags & USRTAU else nlyr + 1
    numu_eff = numu if (flags & USRANG and not onlyfl) else nstr
    bcs = OrcBC()
    keep = []
    for k in ("umu0", "fbeam", "albedo", "btemp", "ttemp", "temis", "fisot", "phi0"):
        v = None if bc is None else bc.get(k)
        if v is not None:
            v = _arr(np.broadcast_to(np.asarray(v, dtype=np.float64), (nwave, ncol)))
            keep.append(v)
            setattr(bcs, k, _p(v))
    temf = _arr(temf)
    wl, wu = _arr(wave_lower), _arr(wave_upper)
    flx = np.zeros((nwave, ncol, ntau_eff, 2))
    uu = None if onlyfl else np.zeros((nwave, ncol, nphi, ntau_eff, numu_eff))
    status = np.zeros((nwave, ncol), dtype=np.int32)
    lib(long_double).orc_forward(
        nwave, ncol, nlyr, nstr, nmom, nprop, flags, accur, ntau, numu, nphi, _p(utau), _p(umu),
        _p(phi), _p(wl), _p(wu), _p(prop), C.byref(bcs), _p(temf), _p(flx), _p(uu),
        status.ctypes.data_as(C.POINTER(C.c_int)), nthreads)
    return flx, uu, status
