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
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "host_tables.hpp"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "disort_solver_body.inc"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "disort_layer_ops.inc"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "disort_adding_step.inc"),
            os.path.join(_D, "..", "..", "pyharp_b200", "csrc", "intensity_device.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hostsim_flux.restype = C.c_int
    lib.hostsim_flux.argtypes = [C.c_int] * 5 + [C.c_uint] + [dp] * 12 + [C.POINTER(C.c_int), dp, dp, C.c_int]
    lib.hostsim_intensity.restype = C.c_int
    lib.hostsim_intensity.argtypes = ([C.c_int] * 6 + [C.c_uint, C.c_double] + [C.c_int] * 3 + [dp] * 17
                                      + [C.POINTER(C.c_int)])
    return lib


_L = None


def flux(prop, bc=None, temf=None, *, nstr, flags, wave_lower=None, wave_upper=None, weights=None, bc_bcast=False,
         layerwise=False):
    """Flux kernel body on the host.  With `weights` (nwave) the BAND instantiation runs instead (fused band
    integration): returns (bflx (ncol, nlyr+1, 2), status).  bc_bcast: boundary fields are (ncol) arrays."""
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
    sh = (ncol,) if bc_bcast else (nwave, ncol)
    bflx_t = None if weights is None else np.zeros((nlyr + 1, 2, ncol))
    _L.hostsim_set_layerwise(int(layerwise))  # the order of the warp-per-solve kernel (nstr <= 8)
    rc = _L.hostsim_flux(nwave, ncol, nlyr, nstr, nprop, flags, P(prop),
                         P(bc.get("umu0"), sh), P(bc.get("fbeam"), sh), P(bc.get("albedo"), sh),
                         P(bc.get("btemp"), sh), P(bc.get("ttemp"), sh), P(bc.get("temis"), sh),
                         P(bc.get("fisot"), sh), P(temf), P(wave_lower), P(wave_upper),
                         flx.ctypes.data_as(dp), st.ctypes.data_as(C.POINTER(C.c_int)), P(weights),
                         None if bflx_t is None else bflx_t.ctypes.data_as(dp), int(bc_bcast))
    assert rc == 0
    if bflx_t is not None:
        return np.ascontiguousarray(bflx_t.transpose(2, 0, 1)), st
    return flx, st


def intensity(prop, bc=None, temf=None, *, nstr, nmom=None, flags, accur=0.0, utau=None, umu=None, phi=None,
              wave_lower=None, wave_upper=None):
    """General path (user levels / angles, azimuth series, NT corrections) on the host simulator.
    Returns (flx, uu, status) shaped like oracle_lib.forward."""
    global _L
    if _L is None:
        _L = _lib()
    prop = np.ascontiguousarray(prop, dtype=np.float64)
    nwave, ncol, nlyr, nprop = prop.shape
    nmom = nstr if nmom is None else nmom
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
    onlyfl = bool(flags & 16)
    ntau = len(utau) if utau is not None else 0
    numu = len(umu) if umu is not None else 0
    nphi = len(phi) if phi is not None else 0
    ntau_eff = ntau if flags & 1 else nlyr + 1
    numu_eff = numu if (flags & 2 and not onlyfl) else nstr
    flx = np.zeros((nwave, ncol, ntau_eff, 2))
    uu = None if onlyfl else np.zeros((nwave, ncol, nphi, ntau_eff, numu_eff))
    st = np.zeros((nwave, ncol), dtype=np.int32)
    sh = (nwave, ncol)
    rc = _L.hostsim_intensity(nwave, ncol, nlyr, nstr, nmom, nprop, flags, accur, ntau, numu, nphi, P(utau), P(umu),
                              P(phi), P(prop), P(bc.get("umu0"), sh), P(bc.get("fbeam"), sh),
                              P(bc.get("albedo"), sh), P(bc.get("btemp"), sh), P(bc.get("ttemp"), sh),
                              P(bc.get("temis"), sh), P(bc.get("fisot"), sh), P(bc.get("phi0"), sh), P(temf),
                              P(wave_lower), P(wave_upper), flx.ctypes.data_as(dp),
                              None if uu is None else uu.ctypes.data_as(dp), st.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return flx, uu, st
