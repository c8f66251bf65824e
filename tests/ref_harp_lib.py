"""ctypes binding of oracle/_ref/libharp_ref.so -- the REFERENCE's own opacity / interpolation /
layer2level code compiled from /root/reference by oracle/ref_harp/Makefile.  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libharp_ref.so")
_dp = C.POINTER(C.c_double)
_L = None


def available():
    return os.path.exists(PATH)


def lib():
    global _L
    if _L is None:
        import torch  # noqa: F401  (loads libtorch so the reference objects resolve their symbols)
        _L = C.CDLL(PATH)
    return _L


def _p(a):
    return a.ctypes.data_as(_dp)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def interpn(coord, data, axes):
    """coord (npoint, ndim); data (*dims, nval); axes = list of 1-D arrays."""
    coord, data = _d(coord), _d(data)
    ndim, nval = coord.shape[1], data.shape[-1]
    dims = np.array(data.shape[:-1], dtype=np.int64)
    axis = _d(np.concatenate(axes))
    out = np.zeros((coord.shape[0], nval))
    rc = lib().ref_interpn(ndim, nval, dims.ctypes.data_as(C.POINTER(C.c_int64)), _p(axis), _p(data), _p(coord),
                           C.c_long(coord.shape[0]), _p(out))
    assert rc == 0
    return out


def aerosol_forward(kind, file, species_weights, species_id, wave, conc, by_wavenumber=False):
    """kind: 's8_fuller' | 'h2so4_simple'.  Returns (out (nwave,ncol,nlyr,2), kwave, kdata)."""
    w, wave, conc = _d(species_weights), _d(wave), _d(conc)
    ncol, nlyr, nsp = conc.shape
    out = np.zeros((len(wave), ncol, nlyr, 2))
    kwave, kdata, rows = np.zeros(4096), np.zeros(8192), C.c_int(0)
    rc = lib().ref_aerosol_forward(0 if kind == "s8_fuller" else 1, file.encode(), _p(w), nsp, species_id, _p(wave),
                                   len(wave), int(by_wavenumber), _p(conc), ncol, nlyr, _p(out), _p(kwave), _p(kdata),
                                   C.byref(rows))
    assert rc == 0
    return out, kwave[:rows.value].copy(), kdata[:2 * rows.value].reshape(-1, 2).copy()


def rfm_forward(kaxis, kdata, krefatm, pres, temp, conc, species_id):
    kaxis, kdata, krefatm, pres, temp, conc = map(_d, (kaxis, kdata, krefatm, pres, temp, conc))
    ncol, nlyr, nsp = conc.shape
    kshape = np.array(kdata.shape, dtype=np.int64)
    out = np.zeros((kdata.shape[0], ncol, nlyr, 1))
    rc = lib().ref_rfm_forward(_p(kaxis), _p(kdata), kshape.ctypes.data_as(C.POINTER(C.c_int64)), _p(krefatm), _p(pres),
                               _p(temp), _p(conc), ncol, nlyr, nsp, species_id, _p(out))
    assert rc == 0
    return out


def layer2level(var, order=4, blower=0, bupper=1):
    var = _d(var)
    out = np.zeros((var.shape[0], var.shape[1] + 1))
    rc = lib().ref_layer2level(_p(var), C.c_long(var.shape[0]), var.shape[1], order, blower, bupper, _p(out))
    assert rc == 0
    return out


def spherical_flux_correction(flx, x1f, area, vol):
    flx, x1f, area, vol = _d(flx).copy(), _d(x1f), _d(area), _d(vol)
    rc = lib().ref_spherical_flux_correction(_p(flx), _p(x1f), _p(area), _p(vol), C.c_long(flx.shape[0]), flx.shape[1])
    assert rc == 0
    return flx
