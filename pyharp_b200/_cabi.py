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


class Attenuator(C.Structure):
    """b200_attenuator (include/b200band.h)."""
    _fields_ = [("type", C.c_int), ("species_id", C.c_int), ("kaxis", _dp), ("kdata", _dp), ("dims", C.c_int64 * 3),
                ("klnp", _dp), ("ktemp", _dp), ("nref", C.c_int), ("kwave", _dp)]


class BandConfig(C.Structure):
    _fields_ = [("disort", Config), ("natt", C.c_int), ("att", C.POINTER(Attenuator)), ("nspecies", C.c_int),
                ("wave", _dp), ("wave_is_wavenumber", C.c_int), ("weights", _dp), ("l2l_order", C.c_int),
                ("l2l_blower", C.c_int), ("l2l_bupper", C.c_int), ("exact_interp", C.c_int), ("mix_eps", C.c_double),
                ("host_chunk_cols", C.c_int)]


class BandIO(C.Structure):
    _fields_ = [("conc", C.c_void_p), ("pres", C.c_void_p), ("temp", C.c_void_p), ("dz", C.c_void_p), ("bc", BC),
                ("bc_per_column", C.c_int), ("rho", C.c_void_p), ("cp", C.c_double), ("bflx", C.c_void_p),
                ("dTdt", C.c_void_p), ("status", C.c_void_p), ("flx", C.c_void_p), ("prop", C.c_void_p)]


ATT_RFM, ATT_AEROSOL = 1, 2

EXPORTS = ("b200disort_create", "b200disort_solve", "b200disort_solve_host",
           "b200disort_check_status", "b200disort_get_rad", "b200disort_reduce_band",
           "b200disort_get_stats", "b200disort_set_stream", "b200disort_set_warp_solve_max", "b200disort_set_layer_split", "b200disort_fp64_peak", "b200disort_destroy",
           "b200disort_strerror", "b200disort_version")
BAND_EXPORTS = ("b200_band_create", "b200_band_nprop", "b200_band_optical_props", "b200_band_forward", "b200_band_forward_host", "b200_band_check",
                "b200_band_allreduce", "b200_band_solver", "b200_band_destroy", "b200disort_allreduce_band",
                "b200disort_solve_band")
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
    # include/b200band.h
    L.b200_band_create.restype = C.c_int
    L.b200_band_create.argtypes = [C.POINTER(BandConfig), C.POINTER(vp)]
    L.b200_band_nprop.restype = C.c_int
    L.b200_band_nprop.argtypes = [vp]
    for name in ("b200_band_forward", "b200_band_forward_host", "b200_band_optical_props"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, C.POINTER(BandIO)]
    L.b200_band_check.restype = C.c_int
    L.b200_band_check.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200_band_allreduce.restype = C.c_int
    L.b200_band_allreduce.argtypes = [vp, vp, C.c_long, vp]
    L.b200_band_solver.restype = vp
    L.b200_band_solver.argtypes = [vp]
    L.b200_band_destroy.restype = None
    L.b200_band_destroy.argtypes = [vp]
    L.b200disort_allreduce_band.restype = C.c_int
    L.b200disort_allreduce_band.argtypes = [vp, vp, vp]
    L.b200disort_solve_band.restype = C.c_int
    L.b200disort_solve_band.argtypes = [vp, vp, C.c_int, C.POINTER(BC), vp, vp, vp, vp]
    L.b200disort_set_layer_split.restype = C.c_int
    L.b200disort_set_layer_split.argtypes = [vp, C.c_int]
    L.b200disort_set_warp_solve_max.restype = C.c_int
    L.b200disort_set_warp_solve_max.argtypes = [vp, C.c_long]
    L.b200disort_set_stream.restype = C.c_int
    L.b200disort_set_stream.argtypes = [vp, vp]
    _lib = L
    return L


def strerror(code):
    return lib().b200disort_strerror(code).decode()
