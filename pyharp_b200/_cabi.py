"""ctypes binding of the product C-ABI library (pyharp_b200/libb200disort.so).

Every symbol declared in include/b200disort.h is bound here.  The library is CUDA-only:
there is no CPU fallback, and a missing library is a hard error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200DISORT_LIB", os.path.join(_HERE, "libb200disort.so"))  # env override: kernel experiments

USRTAU, USRANG, LAMBER, PLANCK, ONLYFL = 1, 2, 4, 8, 16
INTENSITY_CORRECTION, OLD_INTENSITY_CORRECTION, QUIET = 32, 64, 128

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_SOLVE = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)


class Config(C.Structure):
    _fields_ = [("nwave", C.c_int), ("ncol", C.c_int), ("nlyr", C.c_int), ("nstr", C.c_int),
                ("nmom", C.c_int), ("ntau", C.c_int), ("numu", C.c_int), ("nphi", C.c_int),
                ("flags", C.c_uint), ("accur", C.c_double), ("user_tau", _dp), ("user_mu", _dp),
                ("user_phi", _dp), ("wave_lower", _dp), ("wave_upper", _dp), ("device", C.c_int),
                ("stream", C.c_void_p)]


class BC(C.Structure):
    FIELDS = ("umu0", "fbeam", "albedo", "btemp", "ttemp", "temis", "fisot", "phi0")
    _fields_ = [(k, C.c_void_p) for k in FIELDS]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int), ("solves_timed", C.c_int), ("solve_kernel_ms", C.c_float),
                ("total_device_ms", C.c_float), ("workspace_bytes", C.c_uint64)]


EXPORTS = ("b200disort_create", "b200disort_solve", "b200disort_solve_host",
           "b200disort_check_status", "b200disort_get_rad", "b200disort_reduce_band",
           "b200disort_get_stats", "b200disort_fp64_peak", "b200disort_destroy",
           "b200disort_strerror", "b200disort_version")
OPACITY_EXPORTS = ("b200_interpn", "b200_aerosol_optics", "b200_rfm_optics", "b200_layer2level", "b200_band_optics",
                   "b200_spherical_flux_correction", "b200_heating_rate")

_lib = None


def lib():
    """Load the library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  pyharp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.b200disort_create.restype = C.c_int
    L.b200disort_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    for name in ("b200disort_solve", "b200disort_solve_host"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, C.c_int, C.POINTER(BC), vp, vp, vp, vp]
    L.b200disort_check_status.restype = C.c_int
    L.b200disort_check_status.argtypes = [vp, vp, C.POINTER(C.c_int)]
    L.b200disort_get_rad.restype = C.c_int
    L.b200disort_get_rad.argtypes = [vp, vp]
    L.b200disort_reduce_band.restype = C.c_int
    L.b200disort_reduce_band.argtypes = [vp, vp, vp, vp]
    L.b200disort_get_stats.restype = C.c_int
    L.b200disort_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b200disort_fp64_peak.restype = C.c_int
    L.b200disort_fp64_peak.argtypes = [C.c_int, _dp]
    L.b200disort_destroy.restype = None
    L.b200disort_destroy.argtypes = [vp]
    L.b200disort_strerror.restype = C.c_char_p
    L.b200disort_strerror.argtypes = [C.c_int]
    L.b200disort_version.restype = C.c_char_p
    # include/b200opacity.h
    i64p = C.POINTER(C.c_int64)
    L.b200_interpn.restype = C.c_int
    L.b200_interpn.argtypes = [vp, C.c_long, C.c_long, vp, vp, i64p, C.c_int, C.c_int, vp, vp]
    L.b200_aerosol_optics.restype = C.c_int
    L.b200_aerosol_optics.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.b200_rfm_optics.restype = C.c_int
    L.b200_rfm_optics.argtypes = [vp, vp, i64p, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.b200_layer2level.restype = C.c_int
    L.b200_layer2level.argtypes = [vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.b200_spherical_flux_correction.restype = C.c_int
    L.b200_spherical_flux_correction.argtypes = [vp, vp, vp, vp, C.c_long, C.c_int, vp]
    L.b200_heating_rate.restype = C.c_int
    L.b200_heating_rate.argtypes = [vp, vp, vp, C.c_double, C.c_long, C.c_int, vp, vp]
    L.b200_band_optics.restype = C.c_int
    L.b200_band_optics.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   vp, vp]
    _lib = L
    return L


def strerror(code):
    return lib().b200disort_strerror(code).decode()
