"""Host-side mirror of pyharp's opacity plugins and pre-processing on the GPU (SURVEY.md section 8 rows
f-1, f-2, f-3): `AttenuatorOptions`, `S8Fuller`, `H2SO4Simple`, `RFM`, `layer2level`, `band_optics`,
same names / argument meaning / tensor layouts as src/opacity/*.hpp, src/utils/layer2level.hpp and the
mixing in examples/amars_sw.cpp:262-273.  Tensors live on the GPU; every forward() is one kernel of
include/b200opacity.h writing the solver's (nwave, ncol, nlyr, nprop) layout, so
`Disort.forward(band_optics([...]), bc)` runs without a host round-trip.  No CPU path."""
import ctypes as C

import numpy as np
import torch

from . import _cabi

k2ndOrder, k4thOrder = 2, 4
kExtrapolate, kConstant = 0, 1


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): " + ("invalid argument" if rc == 1 else "CUDA launch error"))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, device):
    return torch.as_tensor(t, dtype=torch.float64).to(device).contiguous()


class AttenuatorOptions:
    """src/opacity/attenuator_options.hpp:8-18."""

    def __init__(self):
        self._type, self._opacity_files, self._species_ids = "", [], [0]
        self._species_names, self._species_weights = [], []

    def _acc(name):
        def f(self, v=None):
            if v is None:
                return getattr(self, name)
            setattr(self, name, v)
            return self
        return f

    type = _acc("_type")
    opacity_files = _acc("_opacity_files")
    species_ids = _acc("_species_ids")
    species_names = _acc("_species_names")
    species_weights = _acc("_species_weights")
    del _acc


def read_table(path):
    """decomment + whitespace table, like harp::decomment_file + the row/column scan of S8FullerImpl::reset."""
    rows = []
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if line:
                rows.append([float(x) for x in line.split()])
    return np.array(rows, dtype=np.float64)


class _AerosolTable:
    """S8FullerImpl / H2SO4SimpleImpl (src/opacity/s8_fuller.cpp:10-117, h2so4_simple.cpp:10-117): a
    (wavelength, k_ext [m^2/kg], ssa) table; forward() -> (nwave, ncol, nlyr, 2) = (extinction 1/m,
    ssa-weighted extinction)."""
    TYPE = ""

    def __init__(self, options, table=None, device="cuda"):
        if len(options.species_ids()) != 1 or options.species_ids()[0] < 0:
            raise ValueError("exactly one non-negative species id is required")
        if options.type() not in ("", self.TYPE):
            raise ValueError(f"Mismatch type: {options.type()}")
        self.options = options
        self.device = torch.device(device)
        if table is None:
            if len(options.opacity_files()) != 1:
                raise ValueError("Only one opacity file is allowed")
            table = read_table(options.opacity_files()[0])
        table = np.asarray(table, dtype=np.float64)
        if table.ndim != 2 or table.shape[1] != 3 or table.shape[0] < 1:
            raise ValueError("Invalid table: need rows of (wavelength, k_ext, ssa)")
        kdata = table[:, 1:].copy()
        kdata[:, 0] *= options.species_weights()[options.species_ids()[0]]  # m^2/kg -> m^2/mol
        self.kwave = _dev(table[:, 0], self.device)
        self.kdata = _dev(kdata, self.device)

    def forward(self, conc, kwargs):
        if "wavelength" in kwargs:
            coord, is_wn = _dev(kwargs["wavelength"], self.device), 0
        elif "wavenumber" in kwargs:
            coord, is_wn = _dev(kwargs["wavenumber"], self.device), 1  # 1e4 / wavenumber happens in the kernel
        else:
            raise ValueError("wavelength or wavenumber is required in kwargs")
        conc = _dev(conc, self.device)
        ncol, nlyr, nsp = conc.shape
        out = torch.empty((coord.shape[0], ncol, nlyr, 2), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().b200_aerosol_optics(self.kwave.data_ptr(), self.kdata.data_ptr(), self.kwave.shape[0],
                                                 coord.data_ptr(), coord.shape[0], is_wn, conc.data_ptr(), ncol, nlyr, nsp,
                                                 self.options.species_ids()[0], out.data_ptr(), _stream())
        _check(rc, "b200_aerosol_optics")
        return out

    __call__ = forward


class S8Fuller(_AerosolTable):
    TYPE = "s8_fuller"


class H2SO4Simple(_AerosolTable):
    TYPE = "h2so4_simple"


class RFM:
    """RFMImpl (src/opacity/rfm.cpp:19-169) on an in-memory table: kaxis = [wavenumber | ln p | dT],
    kdata (nwave, npres, ntemp) = ln(m^2/kmol), krefatm (2, npres) = (ln p, T).  The reference reads these
    from NetCDF (absent here); pass arrays instead.  forward() -> (nwave, ncol, nlyr, 1) extinction 1/m."""

    def __init__(self, options, kaxis, kdata, krefatm, device="cuda"):
        self.options = options
        self.device = torch.device(device)
        self.kdata = _dev(kdata, self.device)
        self.kshape = tuple(self.kdata.shape)
        self.kaxis = _dev(kaxis, self.device)
        self.krefatm = _dev(krefatm, self.device)
        if self.kaxis.numel() != sum(self.kshape) or tuple(self.krefatm.shape) != (2, self.kshape[1]):
            raise ValueError("inconsistent RFM table shapes")

    def forward(self, conc, kwargs):
        if "pres" not in kwargs or "temp" not in kwargs:
            raise ValueError("pres and temp are required in kwargs")
        conc = _dev(conc, self.device)
        pres, temp = _dev(kwargs["pres"], self.device), _dev(kwargs["temp"], self.device)
        ncol, nlyr, nsp = conc.shape
        out = torch.empty((self.kshape[0], ncol, nlyr, 1), dtype=torch.float64, device=self.device)
        dims = (C.c_int64 * 3)(*self.kshape)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().b200_rfm_optics(self.kaxis.data_ptr(), self.kdata.data_ptr(), dims,
                                             self.krefatm[0].contiguous().data_ptr(),
                                             self.krefatm[1].contiguous().data_ptr(), self.kshape[1], pres.data_ptr(),
                                             temp.data_ptr(), conc.data_ptr(), ncol, nlyr, nsp,
                                             self.options.species_ids()[0], out.data_ptr(), _stream())
        _check(rc, "b200_rfm_optics")
        return out

    __call__ = forward


def interpn(coord, data, axes, device="cuda"):
    """harp::call_interpn_cuda<N>: coord (npoint, ndim), data (*dims, nval), axes list of 1-D arrays."""
    device = torch.device(device)
    coord, data = _dev(coord, device), _dev(data, device)
    axis = _dev(torch.cat([torch.as_tensor(a, dtype=torch.float64).flatten() for a in axes]), device)
    ndim, nval = coord.shape[1], data.shape[-1]
    dims = (C.c_int64 * ndim)(*data.shape[:-1])
    out = torch.empty((coord.shape[0], nval), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        rc = _cabi.lib().b200_interpn(coord.data_ptr(), coord.shape[0], ndim, data.data_ptr(), axis.data_ptr(), dims,
                                      ndim, nval, out.data_ptr(), _stream())
    _check(rc, "b200_interpn")
    return out


class Layer2LevelOptions:
    """src/utils/layer2level.hpp:22-29."""

    def __init__(self, order=k4thOrder, blower=kExtrapolate, bupper=kConstant, check_positivity=True):
        self.order, self.blower, self.bupper, self.check_positivity = order, blower, bupper, check_positivity


def layer2level(var, options=None):
    """harp::layer2level (src/utils/layer2level.cpp:7-79): (..., nlyr) -> (..., nlyr + 1) on the GPU."""
    options = options or Layer2LevelOptions()
    if options.order not in (k2ndOrder, k4thOrder):
        raise ValueError("Unsupported interpolation order")
    if options.blower not in (kExtrapolate, kConstant) or options.bupper not in (kExtrapolate, kConstant):
        raise ValueError("Unsupported boundary condition")
    if not var.is_cuda:
        raise RuntimeError("layer2level: the B200 path needs a CUDA tensor (no CPU fallback)")
    v = var.to(torch.float64).contiguous()
    nlyr = v.shape[-1]
    rows = v.numel() // nlyr
    out = torch.empty(v.shape[:-1] + (nlyr + 1,), dtype=torch.float64, device=v.device)
    neg = torch.zeros(1, dtype=torch.int32, device=v.device)
    with torch.cuda.device(v.device):
        rc = _cabi.lib().b200_layer2level(v.data_ptr(), rows, nlyr, 1 if options.order == k4thOrder else 0, options.blower,
                                          options.bupper, out.data_ptr(),
                                          neg.data_ptr() if options.check_positivity else None, _stream())
    _check(rc, "b200_layer2level")
    if options.check_positivity and int(neg.item()):
        raise RuntimeError("layer2level check failed: negative values found at cell interface")
    return out


def band_optics(attenuations, dz):
    """Mix attenuator outputs into the solver's input (examples/amars_sw.cpp:262-273):
    prop = sum_a att_a;  prop[..., 0] *= dz;  prop[..., 1] /= summed extinction.  dz: (nlyr,) or (nlyr, 1)."""
    atts = [a.contiguous() for a in attenuations]
    if not atts:
        raise ValueError("band_optics: no attenuator outputs")
    for a in atts:  # raw pointers go to the kernel: every output must be a float64 CUDA tensor of the same leading shape
        if not (a.is_cuda and a.dtype == torch.float64 and a.dim() == 4):
            raise TypeError("band_optics: attenuator outputs must be 4-D float64 CUDA tensors (nwave, ncol, nlyr, nprop)")
        if a.device != atts[0].device or tuple(a.shape[:3]) != tuple(atts[0].shape[:3]):
            raise ValueError("band_optics: attenuator outputs must share their device and (nwave, ncol, nlyr)")
    nwave, ncol, nlyr, _ = atts[0].shape
    nprop = max(a.shape[3] for a in atts)
    dev = atts[0].device
    dz = _dev(dz, dev).reshape(-1)
    if dz.numel() != nlyr:
        raise ValueError("dz must have nlyr entries")
    out = torch.empty((nwave, ncol, nlyr, nprop), dtype=torch.float64, device=dev)
    ptrs = (C.c_void_p * len(atts))(*[a.data_ptr() for a in atts])
    nps = (C.c_int * len(atts))(*[a.shape[3] for a in atts])
    with torch.cuda.device(dev):
        rc = _cabi.lib().b200_band_optics(ptrs, nps, len(atts), dz.data_ptr(), nwave, ncol, nlyr, nprop, out.data_ptr(),
                                          _stream())
    _check(rc, "b200_band_optics")
    return out


def spherical_flux_correction(flx, x1f, area, vol):
    """harp::spherical_flux_correction (src/utils/spherical_flux_correction.cpp:3-18) on the GPU, in place on a
    contiguous CUDA tensor flx (..., nx1); x1f (nx1,), area / vol (..., nx1).  Returns flx."""
    if not (flx.is_cuda and flx.is_contiguous() and flx.dtype == torch.float64):
        raise RuntimeError("spherical_flux_correction: contiguous float64 CUDA tensor required (no CPU fallback)")
    nx1 = x1f.shape[0]
    if flx.shape[-1] != nx1:
        raise ValueError("flx's last dimension must match x1f")
    dev = flx.device
    x1f, area, vol = _dev(x1f, dev), _dev(area, dev).expand_as(flx).contiguous(), _dev(vol, dev).expand_as(flx).contiguous()
    with torch.cuda.device(dev):
        rc = _cabi.lib().b200_spherical_flux_correction(flx.data_ptr(), x1f.data_ptr(), area.data_ptr(), vol.data_ptr(),
                                                        flx.numel() // nx1, nx1, _stream())
    _check(rc, "b200_spherical_flux_correction")
    return flx


def heating_rate(bflx, rho, dz, cp):
    """Heating rate [K/s] per layer from band-integrated fluxes on the GPU, the post-processing of
    examples/amars_sw.cpp:290-301: bflx (..., nlyr+1, 2) = (up, down), level 0 at the surface; rho, dz (..., nlyr).
    dT/dt[k] = -(1 / (rho[k] cp)) (net[k+1] - net[k]) / dz[k] with net = up - down."""
    if not (bflx.is_cuda and bflx.dtype == torch.float64):
        raise RuntimeError("heating_rate: float64 CUDA tensor required (no CPU fallback)")
    bflx = bflx.contiguous()
    nlyr = bflx.shape[-2] - 1
    if bflx.shape[-1] != 2 or nlyr < 1:
        raise ValueError("bflx must be (..., nlyr+1, 2)")
    dev = bflx.device
    lead = bflx.shape[:-2]
    rho = _dev(rho, dev).expand(*lead, nlyr).contiguous()
    dz = _dev(dz, dev).expand(*lead, nlyr).contiguous()
    out = torch.empty(*lead, nlyr, dtype=torch.float64, device=dev)
    ncol = out.numel() // nlyr
    with torch.cuda.device(dev):
        rc = _cabi.lib().b200_heating_rate(bflx.data_ptr(), rho.data_ptr(), dz.data_ptr(), float(cp), ncol, nlyr,
                                           out.data_ptr(), _stream())
    _check(rc, "b200_heating_rate")
    return out
