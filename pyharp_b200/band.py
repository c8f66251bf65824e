"""Host-side mirror of pyharp's `RadiationBandOptions` / `RadiationBand` (src/radiation/radiation_band.hpp:26-89,
radiation_band.cpp:24-128) on top of the band-level C ABI (include/b200band.h).

One `forward` runs the whole per-band pipeline on the GPU -- attenuator optics (RFM / S8Fuller / H2SO4Simple), the
property mixing, layer2level + Planck pre-pass, the DISORT solve of every (spectral point, column) and the band
integration sum_n w_n F_n fused into the solve kernel's flux epilogue -- so that only (ncol, nlyr, .) inputs and the
(ncol, nlyr+1, 2) band flux cross the PCIe bus for CPU tensors.  torch is plumbing (device memory, streams); there
is no CPU path.

    forward(conc, dz, bc, op)        per-spectral-point fluxes (nwave, ncol, nlyr+1, 2): what the reference's
                                     RadiationBandImpl::forward returns today (radiation_band.cpp:123-128)
    forward_band(conc, dz, bc, op)   band-integrated flux (ncol, nlyr+1, 2): the accumulation the reference leaves
                                     commented out (radiation_band.cpp:130-131), examples/amars_lw.cpp:84-88
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .disort import DisortOptions
from .opacity import Layer2LevelOptions, k4thOrder, H2SO4Simple, RFM, S8Fuller

_dp = C.POINTER(C.c_double)


def _host(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class RadiationBandOptions:
    """src/radiation/radiation_band.hpp:26-51 plus the band weights the reference reads with read_weights_rfm
    (src/utils/read_weights.cpp:18-46) or forms from the spectral grid (examples/amars_sw.cpp:185)."""

    def __init__(self):
        self._name, self._outdirs, self._solver_name = "B1", "", "disort"
        self._attenuators = {}
        self._l2l = Layer2LevelOptions()
        self._disort = DisortOptions()
        self._nlyr, self._ncol, self._nmax_prop = 1, 1, 1
        self._wave_lower, self._wave_upper = [], []
        self._weights = None
        self._mix_eps = 0.0
        self._exact_interp = False
        self._host_chunk_cols = 0

    def _acc(name):
        def f(self, v=None):
            if v is None:
                return getattr(self, name)
            setattr(self, name, v)
            return self
        return f

    name = _acc("_name")
    outdirs = _acc("_outdirs")
    solver_name = _acc("_solver_name")
    attenuators = _acc("_attenuators")   # name -> attenuator module (opacity.RFM / S8Fuller / H2SO4Simple)
    l2l = _acc("_l2l")
    disort = _acc("_disort")
    nlyr = _acc("_nlyr")
    ncol = _acc("_ncol")
    nmax_prop = _acc("_nmax_prop")
    wave_lower = _acc("_wave_lower")
    wave_upper = _acc("_wave_upper")
    weights = _acc("_weights")           # (nwave) band weights
    mix_eps = _acc("_mix_eps")           # 0: the examples' mixing; 1e-10: RadiationBandImpl::forward's guard
    exact_interp = _acc("_exact_interp")
    host_chunk_cols = _acc("_host_chunk_cols")  # columns per chunk of the host path (0 = one resident kernel wave)
    del _acc


class RadiationBand:
    """harp::RadiationBand: attenuators + DISORT solver behind one band-level handle."""

    def __init__(self, options, wave=None, wave_is_wavenumber=True, device=None, nspecies=None):
        """wave: the band's spectral grid for aerosol attenuators (the reference passes it per call as
        op['wavenumber'] / op['wavelength'], src/opacity/s8_fuller.cpp:76-86; it is fixed per band).
        nspecies: last dimension of conc (default: what the attenuators' species lists imply)."""
        if options.solver_name() != "disort":
            raise ValueError(f"Unknown solver: {options.solver_name()}")
        if not torch.cuda.is_available():
            raise RuntimeError("pyharp_b200.RadiationBand needs a CUDA device (sm_100a); there is no CPU fallback")
        self.options = options
        self._lib = _cabi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        dop = options.disort()
        nwave = len(options.wave_lower()) or dop.nwave()
        # RadiationBandImpl::reset (radiation_band.cpp:57-66): the band's shape goes into the solver options
        dop.ds().nlyr = options.nlyr()
        dop.nwave(nwave).ncol(options.ncol())
        if len(options.wave_lower()):
            dop.wave_lower(list(options.wave_lower())).wave_upper(list(options.wave_upper()))
        d = dop.ds()
        flags = dop.flag_bits()
        if not (flags & _cabi.ONLYFL) or (flags & _cabi.USRTAU):
            raise NotImplementedError("RadiationBand drives the flux path: the disort flags need 'onlyfl' and no 'usrtau'")
        self._flags = flags
        self.nwave, self.ncol, self.nlyr = nwave, options.ncol(), options.nlyr()
        w = options.weights()
        if w is None:
            w = np.ones(nwave)
        self._keep = []

        def ptr(a):
            a = _host(a)
            self._keep.append(a)
            return a.ctypes.data_as(_dp)

        atts = list(options.attenuators().items())  # std::map order = sorted keys; Python dicts keep insertion order
        if not atts:
            raise ValueError("a band needs at least one attenuator")
        arr = (_cabi.Attenuator * len(atts))()
        nspecies_arg, nspecies = nspecies, 0
        for k, (_, m) in enumerate(atts):
            a = arr[k]
            a.species_id = m.options.species_ids()[0]
            nspecies = max(nspecies, a.species_id + 1, len(m.options.species_names()), len(m.options.species_weights()))
            if isinstance(m, RFM):
                a.type = _cabi.ATT_RFM
                a.kaxis, a.kdata = ptr(m.kaxis), ptr(m.kdata)
                for i in range(3):
                    a.dims[i] = m.kshape[i]
                a.klnp, a.ktemp, a.nref = ptr(m.krefatm[0]), ptr(m.krefatm[1]), m.kshape[1]
            elif isinstance(m, (S8Fuller, H2SO4Simple)):
                a.type = _cabi.ATT_AEROSOL
                a.kwave, a.kdata = ptr(m.kwave), ptr(m.kdata)
                a.dims[0] = m.kwave.shape[0]
            else:
                raise TypeError(f"Unknown attenuator type: {type(m).__name__}")
        self.nspecies = nspecies = int(nspecies_arg) if nspecies_arg else nspecies
        cfg = _cabi.BandConfig()
        dc = cfg.disort
        dc.nwave, dc.ncol, dc.nlyr = nwave, self.ncol, self.nlyr
        dc.nstr, dc.nmom = d.nstr, max(d.nmom, d.nstr)
        dc.ntau = dc.numu = dc.nphi = 1
        dc.flags, dc.accur = flags, d.accur
        if len(dop.wave_lower()):
            dc.wave_lower, dc.wave_upper = ptr(dop.wave_lower()), ptr(dop.wave_upper())
        if (flags & _cabi.PLANCK) and len(dop.wave_lower()) != nwave:
            raise ValueError("planck needs wave_lower/wave_upper of length nwave")
        dc.device = self.device.index
        cfg.natt, cfg.att, cfg.nspecies = len(atts), arr, nspecies
        if wave is not None:
            cfg.wave, cfg.wave_is_wavenumber = ptr(wave), int(bool(wave_is_wavenumber))
        cfg.weights = ptr(w)
        l2l = options.l2l()
        cfg.l2l_order = 1 if l2l.order == k4thOrder else 0
        cfg.l2l_blower, cfg.l2l_bupper = l2l.blower, l2l.bupper
        cfg.exact_interp, cfg.mix_eps = int(bool(options.exact_interp())), float(options.mix_eps())
        cfg.host_chunk_cols = int(options.host_chunk_cols())
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream()
            dc.stream = self._stream.cuda_stream
            h = C.c_void_p()
            rc = self._lib.b200_band_create(C.byref(cfg), C.byref(h))
        if rc != _cabi.OK:
            raise RuntimeError(f"b200_band_create failed: {_cabi.strerror(rc)}")
        self._h = h
        self.nprop = self._lib.b200_band_nprop(h)
        options.nmax_prop(self.nprop)
        self.shared = {}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.b200_band_destroy(h)
            self._h = None

    def _use_current_stream(self):
        """Launch on torch's current stream of the band's device (call inside torch.cuda.device(self.device))."""
        st = torch.cuda.current_stream()
        if st.cuda_stream != self._stream.cuda_stream:
            rc = self._lib.b200disort_set_stream(self._lib.b200_band_solver(self._h), C.c_void_p(st.cuda_stream))
            if rc != _cabi.OK:
                raise RuntimeError(f"b200disort_set_stream failed: {_cabi.strerror(rc)}")
            self._stream = st

    # ---- forward -----------------------------------------------------------------------------------------------
    def _io(self, conc, dz, bc, op, dev, rho=None, cp=None):
        nw, nc, nl = self.nwave, self.ncol, self.nlyr
        if tuple(conc.shape[:2]) != (nc, nl):
            raise ValueError(f"conc must be (ncol={nc}, nlyr={nl}, nspecies), got {tuple(conc.shape)}")
        keep = []

        def prep(t, shape=None):
            t = torch.as_tensor(t, dtype=torch.float64).to(dev).contiguous()
            if shape is not None and tuple(t.shape) != shape:
                raise ValueError(f"expected shape {shape}, got {tuple(t.shape)}")
            keep.append(t)
            return t.data_ptr()

        io = _cabi.BandIO()
        conc = torch.as_tensor(conc, dtype=torch.float64).to(dev).contiguous()
        keep.append(conc)
        io.conc = conc.data_ptr()
        if conc.shape[2] != self.nspecies:
            raise ValueError(f"conc must have exactly {self.nspecies} species (the band was created for that count)")
        op = op or {}
        if "pres" in op:
            io.pres = prep(op["pres"], (nc, nl))
        if "temp" in op:
            io.temp = prep(op["temp"], (nc, nl))
        if dz is not None:
            io.dz = prep(torch.as_tensor(dz, dtype=torch.float64).reshape(-1), (nl,))
        per_col = None
        for k, v in (bc or {}).items():
            if k == "fluor":
                continue
            if k not in _cabi.BC.FIELDS or k == "phi0":
                raise KeyError(f"unknown boundary condition '{k}'")
            v = torch.as_tensor(v, dtype=torch.float64)
            pc = v.dim() == 1
            if per_col is None:
                per_col = pc
            if pc != per_col:
                raise ValueError("boundary fields must all be (nwave, ncol) or all be (ncol)")
            setattr(io.bc, k, prep(v, (nc,) if pc else (nw, nc)))
        io.bc_per_column = int(bool(per_col))
        if rho is not None:
            io.rho, io.cp = prep(rho, (nc, nl)), float(cp)
        return io, keep

    def _raise(self, rc, status=None):
        if rc == _cabi.ERR_SOLVE:
            msg = "disort: one or more solves of the band failed (bad input, or a negative interface temperature)"
            if status is not None:
                bad = status.flatten().nonzero().flatten()
                if len(bad):
                    msg = (f"disort: {len(bad)} of {status.numel()} solves failed (first: solve {int(bad[0])}, "
                           f"status {int(status.flatten()[int(bad[0])])})")
            raise RuntimeError(msg)
        if rc != _cabi.OK:
            raise RuntimeError(f"b200 band forward failed: {_cabi.strerror(rc)}")

    def forward_band(self, conc, dz, bc, op, rho=None, cp=None, want_status=False, check=True):
        """Band-integrated flux (ncol, nlyr+1, 2) [and the heating rate (ncol, nlyr) when rho / cp are given].
        CPU tensors take the host entry point (chunked, copies overlapped with the kernels, result in pinned
        memory); CUDA tensors the device entry point (asynchronous; `check` synchronises and raises on failure)."""
        conc = torch.as_tensor(conc)
        on_host = not conc.is_cuda
        dev = torch.device("cpu") if on_host else self.device
        io, keep = self._io(conc, dz, bc, op, dev, rho, cp)
        pin = dict(pin_memory=True) if on_host else {}
        bflx = torch.empty((self.ncol, self.nlyr + 1, 2), dtype=torch.float64, device=dev, **pin)
        io.bflx = bflx.data_ptr()
        dT = status = None
        if rho is not None:
            dT = torch.empty((self.ncol, self.nlyr), dtype=torch.float64, device=dev, **pin)
            io.dTdt = dT.data_ptr()
        if want_status:
            status = torch.empty((self.nwave, self.ncol), dtype=torch.int32, device=dev, **pin)
            io.status = status.data_ptr()
        with torch.cuda.device(self.device):
            self._use_current_stream()
            if on_host:
                rc = self._lib.b200_band_forward_host(self._h, C.byref(io))
            else:
                rc = self._lib.b200_band_forward(self._h, C.byref(io))
                if rc == _cabi.OK and check:
                    rc = self._lib.b200_band_check(self._h, None)
        self._status = status
        self._raise(rc, status if (status is not None and (on_host or check)) else None)
        return (bflx, dT) if rho is not None else bflx

    def forward(self, conc, dz, bc, op):
        """RadiationBandImpl::forward as the reference has it today: the per-spectral-point fluxes
        (nwave, ncol, nlyr+1, 2); the mixed optical properties are exported as
        self.shared['radiation/<name>/optics'] (radiation_band.cpp:119-120) in the solver's 4-D layout.
        CPU tensors are moved to the device first (the reference's callers pass CPU tensors)."""
        conc = torch.as_tensor(conc)
        io, keep = self._io(conc, dz, bc, op, self.device)
        flx = torch.empty((self.nwave, self.ncol, self.nlyr + 1, 2), dtype=torch.float64, device=self.device)
        bflx = torch.empty((self.ncol, self.nlyr + 1, 2), dtype=torch.float64, device=self.device)
        prop = torch.empty((self.nwave, self.ncol, self.nlyr, self.nprop), dtype=torch.float64, device=self.device)
        status = torch.empty((self.nwave, self.ncol), dtype=torch.int32, device=self.device)
        io.bflx, io.flx, io.prop, io.status = bflx.data_ptr(), flx.data_ptr(), prop.data_ptr(), status.data_ptr()
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200_band_forward(self._h, C.byref(io))
            if rc == _cabi.OK:
                rc = self._lib.b200_band_check(self._h, None)
        self._status = status
        self._raise(rc, status)
        self.shared["radiation/" + self.options.name() + "/optics"] = prop
        self.bflx = bflx
        return flx if conc.is_cuda else flx.cpu()

    __call__ = forward

    def optics(self, conc, dz, op):
        """The optics stage alone: prop (nwave, ncol, nlyr, nprop) on the device (b200_band_optical_props)."""
        io, keep = self._io(torch.as_tensor(conc), dz, None, op, self.device)
        prop = torch.empty((self.nwave, self.ncol, self.nlyr, self.nprop), dtype=torch.float64, device=self.device)
        io.prop = prop.data_ptr()
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200_band_optical_props(self._h, C.byref(io))
        self._raise(rc)
        return prop

    def status(self):
        return self._status

    def allreduce(self, bflx, comm):
        """Sum a device band flux over the ranks of an NCCL communicator (`comm`: ncclComm_t as an int, e.g. from
        pyharp_b200.nccl.Communicator) -- the one collective of spectral-point sharding."""
        if not bflx.is_cuda or not bflx.is_contiguous():
            raise ValueError("allreduce: contiguous CUDA tensor required")
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200_band_allreduce(self._h, bflx.data_ptr(), bflx.numel(), C.c_void_p(int(comm)))
        if rc != _cabi.OK:
            raise RuntimeError(f"b200_band_allreduce failed: {_cabi.strerror(rc)}")
        return bflx

    def stats(self):
        st = _cabi.Stats()
        rc = self._lib.b200disort_get_stats(self._lib.b200_band_solver(self._h), C.byref(st))
        if rc != _cabi.OK:
            raise RuntimeError(_cabi.strerror(rc))
        return {"kernel_launches": st.kernel_launches, "solves_timed": st.solves_timed, "solve_kernel_ms": st.solve_kernel_ms,
                "total_device_ms": st.total_device_ms, "workspace_bytes": st.workspace_bytes}
