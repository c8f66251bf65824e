"""Host-side mirror of pydisort's `disort::DisortOptions` / `disort::Disort` as pyharp uses
them (tests/test_disort.cpp:13-55, examples/amars_sw.cpp:43-65,216,280,
examples/amars_lw.cpp:18-38,60,80, src/radiation/radiation_band.cpp:58-68), on top of the
C ABI in include/b200disort.h.  Same names, argument meaning and error behaviour
(exceptions on failure, cf. src/rtsolver/rt_solver_disort.cpp_:147-151).

torch is plumbing only (device memory, streams); all arithmetic happens in the CUDA
library.  There is no CPU path."""
import ctypes as C

import numpy as np
import torch

from . import _cabi

# disort::index (src/index.h:12-18 has the same constants on the pyharp side)
IEX, ISS, IPM = 0, 1, 2
IUP, IDN = 0, 1

kIsotropic, kRayleigh, kHenyeyGreenstein, kHazeGarciaSiewert, kCloudGarciaSiewert = range(5)

_FLAG_BITS = {
    "usrtau": _cabi.USRTAU, "usrang": _cabi.USRANG, "lamber": _cabi.LAMBER,
    "planck": _cabi.PLANCK, "onlyfl": _cabi.ONLYFL,
    "intensity_correction": _cabi.INTENSITY_CORRECTION,
    "old_intensity_correction": _cabi.OLD_INTENSITY_CORRECTION, "quiet": _cabi.QUIET,
}
# tokens cdisort accepts that do not change the numerical path here (printing), and tokens
# for features outside pyharp's path (rejected at construction when set).
_PRINT_TOKENS = {"print-input", "print-fluxes", "print-intensity", "print-transmissivity",
                 "print-phase-function"}
_UNSUPPORTED_TOKENS = {"ibcnd", "spher", "general_source", "output_uum", "brdf"}


class PhaseMomentOptions:
    def __init__(self):
        self._type, self._gg = kIsotropic, 0.0

    def type(self, v=None):
        if v is None:
            return self._type
        self._type = v
        return self

    def gg(self, v=None):
        if v is None:
            return self._gg
        self._gg = float(v)
        return self


def scattering_moments(nmom, opts=None):
    """disort::scattering_moments (tests/test_disort.cpp:37-40): Legendre moments 1..nmom
    (the 0th moment is implicit = 1) as a float64 tensor of shape (nmom,)."""
    opts = opts or PhaseMomentOptions()
    k = torch.arange(1, nmom + 1, dtype=torch.float64)
    t = opts.type()
    if t == kIsotropic:
        return torch.zeros(nmom, dtype=torch.float64)
    if t == kRayleigh:
        out = torch.zeros(nmom, dtype=torch.float64)
        if nmom >= 2:
            out[1] = 0.1
        return out
    if t == kHenyeyGreenstein:
        g = opts.gg()
        if not (-1.0 < g < 1.0):
            raise ValueError("scattering_moments: |gg| must be < 1")
        return g ** k
    raise NotImplementedError("tabulated Garcia-Siewert phase functions are not on pyharp's path")


class DisortState:
    """The fields of cdisort's disort_state that pyharp touches through op.ds()."""

    def __init__(self):
        self.nlyr = 1
        self.nstr = 4
        self.nmom = 4
        self.nphi = 1
        self.ntau = 1
        self.numu = 1
        self.accur = 0.0
        self.utau = []
        self.umu = []
        self.phi = []


class DisortOptions:
    """Fluent option bag: op.nwave(10) sets and returns self, op.nwave() reads."""

    def __init__(self):
        self._header = ""
        self._flags = ""
        self._nwave = 1
        self._ncol = 1
        self._wave_lower = []
        self._wave_upper = []
        self._user_mu = []
        self._user_phi = []
        self._user_tau = []
        self._ds = DisortState()

    def _acc(name):
        def f(self, v=None):
            if v is None:
                return getattr(self, name)
            setattr(self, name, v)
            return self
        return f

    header = _acc("_header")
    flags = _acc("_flags")
    nwave = _acc("_nwave")
    ncol = _acc("_ncol")
    wave_lower = _acc("_wave_lower")
    wave_upper = _acc("_wave_upper")
    user_mu = _acc("_user_mu")
    user_phi = _acc("_user_phi")
    user_tau = _acc("_user_tau")
    del _acc

    def ds(self):
        return self._ds

    def flag_bits(self):
        bits = 0
        for tok in filter(None, (t.strip() for t in self._flags.split(","))):
            if tok in _FLAG_BITS:
                bits |= _FLAG_BITS[tok]
            elif tok in _PRINT_TOKENS:
                continue
            elif tok in _UNSUPPORTED_TOKENS:
                raise NotImplementedError(f"disort flag '{tok}' is outside the B200 solve path")
            else:
                raise ValueError(f"unknown disort flag '{tok}'")
        return bits

    def __str__(self):
        d = self._ds
        return (f"DisortOptions(header={self._header!r}, flags={self._flags!r}, nwave={self._nwave}, "
                f"ncol={self._ncol}, nlyr={d.nlyr}, nstr={d.nstr}, nmom={d.nmom}, "
                f"user_tau={list(self._user_tau)}, user_mu={list(self._user_mu)}, "
                f"user_phi={list(self._user_phi)}, wave=[{list(self._wave_lower)}, {list(self._wave_upper)}])")


def _darr(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Disort:
    """disort::Disort: owns one C-ABI handle (device workspace + tables) for a fixed
    (nwave, ncol, nlyr, nstr) shape.  forward() returns flux (nwave, ncol, nlev, 2)."""

    def __init__(self, options, device=None):
        self.options = options
        self._lib = _cabi.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("pyharp_b200.Disort needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = options.ds()
        flags = options.flag_bits()
        self._flags = flags
        if len(options.user_tau()):
            d.ntau = len(options.user_tau())
            d.utau = list(options.user_tau())
        if len(options.user_mu()):
            d.numu = len(options.user_mu())
            d.umu = list(options.user_mu())
        if len(options.user_phi()):
            d.nphi = len(options.user_phi())
            d.phi = list(options.user_phi())
        self._keep = []
        cfg = _cabi.Config()
        cfg.nwave, cfg.ncol, cfg.nlyr = options.nwave(), options.ncol(), d.nlyr
        cfg.nstr, cfg.nmom = d.nstr, max(d.nmom, d.nstr)
        cfg.ntau, cfg.numu, cfg.nphi = d.ntau, d.numu, d.nphi
        cfg.flags, cfg.accur = flags, d.accur
        for name, vals in (("user_tau", options.user_tau()), ("user_mu", options.user_mu()),
                           ("user_phi", options.user_phi()), ("wave_lower", options.wave_lower()),
                           ("wave_upper", options.wave_upper())):
            if len(vals):
                a, p = _darr(vals)
                self._keep.append(a)
                setattr(cfg, name, p)
        if (flags & _cabi.PLANCK) and len(options.wave_lower()) != options.nwave():
            raise ValueError("planck needs wave_lower/wave_upper of length nwave")
        if self.device.index is None:  # device='cuda': the caller's current device, not GPU 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        cfg.device = self.device.index
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream()
            cfg.stream = self._stream.cuda_stream
            h = C.c_void_p()
            rc = self._lib.b200disort_create(C.byref(cfg), C.byref(h))
        if rc != _cabi.OK:
            raise RuntimeError(f"b200disort_create failed: {_cabi.strerror(rc)}")
        self._h = h
        self._status = None
        self._rad = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.b200disort_destroy(h)
            self._h = None

    def set_warp_solve_max(self, max_solves):
        """Largest flux-only batch that runs on the warp-per-solve form of the kernel (-1: default, 0: never)."""
        rc = self._lib.b200disort_set_warp_solve_max(self._h, int(max_solves))
        if rc != _cabi.OK:
            raise RuntimeError(_cabi.strerror(rc))
        return self

    def set_layer_split(self, mode):
        """nstr = 4, 8, >= 12: 1 = layer-split form of the flux solve, 0 = never, -1 = default (from nstr = 16 up, and at nstr = 8
        for batches of 1185..8192 solves); include/b200disort.h."""
        rc = self._lib.b200disort_set_layer_split(self._h, int(mode))
        if rc != _cabi.OK:
            raise RuntimeError(_cabi.strerror(rc))
        return self

    def ds(self, n=0, j=0):
        return self.options.ds()

    @property
    def nlev(self):
        d = self.options.ds()
        return d.ntau if self._flags & _cabi.USRTAU else d.nlyr + 1

    def _use_current_stream(self):
        """Launch on torch's CURRENT stream of the handle's device (call inside torch.cuda.device(self.device)): the
        outputs are allocated by the caching allocator for that stream and the inputs were produced on it."""
        st = torch.cuda.current_stream()
        if st.cuda_stream != self._stream.cuda_stream:
            rc = self._lib.b200disort_set_stream(self._h, C.c_void_p(st.cuda_stream))
            if rc != _cabi.OK:
                raise RuntimeError(f"b200disort_set_stream failed: {_cabi.strerror(rc)}")
            self._stream = st

    def _check_shapes(self, prop, bc, temf):
        d = self.options.ds()
        nw, nc = self.options.nwave(), self.options.ncol()
        if prop.dim() != 4 or tuple(prop.shape[:3]) != (nw, nc, d.nlyr):
            raise ValueError(f"prop must be (nwave={nw}, ncol={nc}, nlyr={d.nlyr}, nprop), got {tuple(prop.shape)}")
        if prop.dtype != torch.float64:
            raise TypeError("prop must be float64")
        for k, v in (bc or {}).items():
            if k == "fluor":  # accepted and unused, as in cdisort without general_source (the C++ shim does the same)
                continue
            if k not in _cabi.BC.FIELDS:
                raise KeyError(f"unknown boundary condition '{k}'")
            if tuple(v.shape) != (nw, nc):
                raise ValueError(f"bc['{k}'] must be (nwave, ncol)")
        if self._flags & _cabi.PLANCK:
            if temf is None:
                raise ValueError("planck flag set but temf not given")
            if tuple(temf.shape) != (nc, d.nlyr + 1):
                raise ValueError("temf must be (ncol, nlyr+1)")

    def forward(self, prop, bc=None, temf=None):
        """disort::DisortImpl::forward(prop, &bc, temf).  CPU tensors take the host-pointer
        C-ABI entry (pinned staging, H2D/D2H inside); CUDA tensors the device-pointer one."""
        self._check_shapes(prop, bc, temf)
        bc = bc or {}
        nw, nc = self.options.nwave(), self.options.ncol()
        onlyfl = bool(self._flags & _cabi.ONLYFL)
        d = self.options.ds()
        on_host = not prop.is_cuda
        dev = torch.device("cpu") if on_host else self.device
        prop = prop.contiguous()
        bcs = _cabi.BC()
        keep = [prop]
        for k, v in bc.items():
            if k == "fluor":
                continue
            v = v.to(device=dev, dtype=torch.float64).contiguous()
            keep.append(v)
            setattr(bcs, k, v.data_ptr())
        tptr = None
        if temf is not None:
            temf = temf.to(device=dev, dtype=torch.float64).contiguous()
            keep.append(temf)
            tptr = temf.data_ptr()
        # host results land in pinned memory (torch's caching host allocator) so the chunked D2H
        # copies of b200disort_solve_host are true async DMA
        pin = dict(pin_memory=True) if on_host else {}
        flx = torch.empty((nw, nc, self.nlev, 2), dtype=torch.float64, device=dev, **pin)
        status = torch.empty((nw, nc), dtype=torch.int32, device=dev, **pin)
        uu = None
        if not onlyfl:
            numu = d.numu if self._flags & _cabi.USRANG else d.nstr
            uu = torch.empty((nw, nc, d.nphi, self.nlev, numu), dtype=torch.float64, device=dev)
        with torch.cuda.device(self.device):
            self._use_current_stream()
            fn = self._lib.b200disort_solve_host if on_host else self._lib.b200disort_solve
            rc = fn(self._h, prop.data_ptr(), prop.shape[3], C.byref(bcs), tptr, flx.data_ptr(),
                    None if uu is None else uu.data_ptr(), status.data_ptr())
            if rc == _cabi.OK and not on_host:
                nbad = C.c_int(0)
                rc = self._lib.b200disort_check_status(self._h, status.data_ptr(), C.byref(nbad))
        self._status, self._rad = status, uu
        if rc == _cabi.ERR_SOLVE:
            bad = status.flatten().nonzero().flatten()
            first = int(bad[0]) if len(bad) else -1
            code = int(status.flatten()[first]) if len(bad) else -1
            raise RuntimeError(f"disort: {len(bad)} of {nw * nc} solves failed "
                               f"(first: solve {first}, status {code})")
        if rc != _cabi.OK:
            raise RuntimeError(f"b200disort solve failed: {_cabi.strerror(rc)}")
        return flx

    __call__ = forward

    def forward_device_async(self, prop, bcs, temf_ptr, flx, status):
        """Raw device-pointer solve without synchronisation (used by bench.py's kernel-only leg)."""
        self._use_current_stream()
        rc = self._lib.b200disort_solve(self._h, prop.data_ptr(), prop.shape[3], C.byref(bcs), temf_ptr,
                                        flx.data_ptr(), None, status.data_ptr())
        if rc != _cabi.OK:
            raise RuntimeError(f"b200disort_solve failed: {_cabi.strerror(rc)}")

    def forward_band(self, prop, bc, temf, weights, check=True):
        """Solve + band integration in one kernel (b200disort_solve_band): bflx (ncol, nlyr+1, 2) =
        sum_n weights[n] F_n (examples/amars_lw.cpp:84-88) without ever writing the per-spectral-point fluxes.
        CUDA tensors only."""
        self._check_shapes(prop, bc, temf)
        if not prop.is_cuda:
            raise RuntimeError("forward_band: CUDA tensors required (use RadiationBand for host inputs)")
        nw, nc = self.options.nwave(), self.options.ncol()
        prop = prop.contiguous()
        bcs = _cabi.BC()
        keep = [prop]
        for k, v in (bc or {}).items():
            if k == "fluor":
                continue
            v = v.to(device=self.device, dtype=torch.float64).contiguous()
            keep.append(v)
            setattr(bcs, k, v.data_ptr())
        tptr = None
        if temf is not None:
            temf = temf.to(device=self.device, dtype=torch.float64).contiguous()
            keep.append(temf)
            tptr = temf.data_ptr()
        w = weights.to(device=self.device, dtype=torch.float64).contiguous()
        bflx = torch.empty((nc, self.options.ds().nlyr + 1, 2), dtype=torch.float64, device=self.device)
        status = torch.empty((nw, nc), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200disort_solve_band(self._h, prop.data_ptr(), prop.shape[3], C.byref(bcs), tptr, w.data_ptr(),
                                                 bflx.data_ptr(), status.data_ptr())
            if rc == _cabi.OK and check:
                nbad = C.c_int(0)
                rc = self._lib.b200disort_check_status(self._h, status.data_ptr(), C.byref(nbad))
        self._status = status
        if rc == _cabi.ERR_SOLVE:
            raise RuntimeError(f"disort: {int((status != 0).sum())} of {nw * nc} solves failed")
        if rc != _cabi.OK:
            raise RuntimeError(f"b200disort_solve_band failed: {_cabi.strerror(rc)}")
        return bflx

    def allreduce_band(self, bflx, comm):
        """In-place sum of a device band flux (ncol, nlev, 2) over the ranks of an NCCL communicator
        (b200disort_allreduce_band; `comm`: ncclComm_t as an int, e.g. pyharp_b200.nccl.Communicator) -- the one
        collective of spectral-point sharding (examples/amars_lw.cpp:88 across ranks)."""
        if not (bflx.is_cuda and bflx.is_contiguous() and bflx.dtype == torch.float64):
            raise ValueError("allreduce_band: contiguous float64 CUDA tensor required")
        if bflx.numel() != self.options.ncol() * self.nlev * 2:
            raise ValueError("allreduce_band: bflx must be (ncol, nlev, 2)")
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200disort_allreduce_band(self._h, bflx.data_ptr(), C.c_void_p(int(comm)))
        if rc != _cabi.OK:
            raise RuntimeError(f"b200disort_allreduce_band failed: {_cabi.strerror(rc)}")
        return bflx

    def get_rad(self, options=None):
        """Intensities of the last forward(), shape (nwave, ncol, nphi, ntau, numu): cdisort's `uu` memory order
        (numu fastest) behind the (wave, column) batch axes.  pyharp only prints this tensor
        (tests/test_disort.cpp:52), so the reference does not pin the axis order; SURVEY.md 8 row a4 records it as
        inferred from cdisort's layout, and include/b200disort.h documents the same order for the C ABI."""
        if self._rad is None:
            raise RuntimeError("get_rad: no intensities available (onlyfl set or forward not called)")
        return self._rad

    def status(self):
        return self._status

    def stats(self):
        st = _cabi.Stats()
        rc = self._lib.b200disort_get_stats(self._h, C.byref(st))
        if rc != _cabi.OK:
            raise RuntimeError(_cabi.strerror(rc))
        return {"kernel_launches": st.kernel_launches, "solves_timed": st.solves_timed, "solve_kernel_ms": st.solve_kernel_ms,
                "total_device_ms": st.total_device_ms, "workspace_bytes": st.workspace_bytes}

    def reduce_band(self, flx, weights):
        """bflx(ncol, nlev, 2) = sum_n weights[n] flx[n] on the device (examples/amars_lw.cpp:88)."""
        if not flx.is_cuda or flx.device != self.device:
            raise ValueError("reduce_band: flx must be a CUDA tensor on the solver's device")
        w = weights.to(device=self.device, dtype=torch.float64).contiguous()
        out = torch.empty(flx.shape[1:], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._use_current_stream()
            rc = self._lib.b200disort_reduce_band(self._h, flx.contiguous().data_ptr(), w.data_ptr(), out.data_ptr())
        if rc != _cabi.OK:
            raise RuntimeError(_cabi.strerror(rc))
        return out


def fp64_peak_tflops(device=0):
    v = C.c_double(0.0)
    rc = _cabi.lib().b200disort_fp64_peak(device, C.byref(v))
    if rc != _cabi.OK:
        raise RuntimeError(_cabi.strerror(rc))
    return v.value
