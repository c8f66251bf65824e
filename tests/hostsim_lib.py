"""ctypes binding of the test-only host simulator (tests/hostsim/libhostsim.so): the device
math of pyharp_b200/csrc/disort_device.cuh compiled by g++ so the algorithm can be checked
against the oracle on a machine without a GPU.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


def _lib():
    so = os.path.join(_D, "libhostsim.so")
    srcs = [os.path.join(_D, "hostsim.cpp"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "disort_device.cuh"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "host_tables.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hostsim_flux.restype = C.c_int
    lib.hostsim_flux.argtypes = [C.c_int] * 5 + [C.c_uint] + [dp] * 12 + [C.POINTER(C.c_int)]
    return lib


_L = None


def flux(prop, bc=None, temf=None, *, nstr, flags, wave_lower=None, wave_upper=None):
    global _L
    if _L is None:
        _L = _lib()
    prop = np.ascontiguousarray(prop, dtype=np.float64)
    nwave, ncol, nlyr, nprop = prop.shape
    dp = C.POINTER(C.c_double)
    keep = []

    def P(a, shape=None):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float64)
        if shape is not None:
            a = np.broadcast_to(a, shape)
        a = np.ascontiguousarray(a)
        keep.append(a)
        return a.ctypes.data_as(dp)

    bc = bc or {}
    flx = np.zeros((nwave, ncol, nlyr + 1, 2))
    st = np.zeros((nwave, ncol), dtype=np.int32)
    sh = (nwave, ncol)
    rc = _L.hostsim_flux(nwave, ncol, nlyr, nstr, nprop, flags, P(prop),
                         P(bc.get("umu0"), sh), P(bc.get("fbeam"), sh), P(bc.get("albedo"), sh),
                         P(bc.get("btemp"), sh), P(bc.get("ttemp"), sh), P(bc.get("temis"), sh),
                         P(bc.get("fisot"), sh), P(temf), P(wave_lower), P(wave_upper),
                         flx.ctypes.data_as(dp), st.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return flx, st
