"""Host-side mirror of pydisort's option surface as pyharp uses it (tests/test_disort.cpp:13-40,
examples/amars_sw.cpp:43-65, examples/amars_lw.cpp:18-38)."""
import numpy as np
import pytest
import torch

from pyharp_b200 import (DisortOptions, PhaseMomentOptions, scattering_moments, IEX, ISS, IPM, IUP, IDN,
                         kIsotropic, kHenyeyGreenstein, kRayleigh, _cabi, workloads as W)


def test_index_constants():
    # src/index.h:12-18
    assert (IEX, ISS, IPM, IUP, IDN) == (0, 1, 2, 0, 1)


def test_options_like_test_disort_cpp():
    op = DisortOptions()
    op.header("running disort example")
    op.flags("usrtau,usrang,lamber,quiet,intensity_correction,old_intensity_correction,print-input,print-phase-function")
    op.nwave(10)
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = 1, 16, 16
    op.user_mu([-1, -0.5, -0.1, 0.1, 0.5, 1]).user_phi([0]).user_tau([0, 0.03125])
    bits = op.flag_bits()
    assert bits == (_cabi.USRTAU | _cabi.USRANG | _cabi.LAMBER | _cabi.QUIET | _cabi.INTENSITY_CORRECTION
                    | _cabi.OLD_INTENSITY_CORRECTION)
    assert op.nwave() == 10 and op.ncol() == 1
    assert "nstr=16" in str(op)


def test_flag_errors():
    with pytest.raises(ValueError):
        DisortOptions().flags("lamber,bogus").flag_bits()
    with pytest.raises(NotImplementedError):
        DisortOptions().flags("lamber,spher").flag_bits()
    assert DisortOptions().flags("").flag_bits() == 0
    assert DisortOptions().flags(" lamber , planck ").flag_bits() == _cabi.LAMBER | _cabi.PLANCK


def test_scattering_moments():
    assert torch.equal(scattering_moments(16, PhaseMomentOptions().type(kIsotropic)), torch.zeros(16, dtype=torch.float64))
    m = scattering_moments(4, PhaseMomentOptions().type(kHenyeyGreenstein).gg(0.5))
    assert torch.allclose(m, torch.tensor([0.5, 0.25, 0.125, 0.0625], dtype=torch.float64))
    r = scattering_moments(4, PhaseMomentOptions().type(kRayleigh))
    assert r.tolist() == [0.0, 0.1, 0.0, 0.0]
    with pytest.raises(ValueError):
        scattering_moments(4, PhaseMomentOptions().type(kHenyeyGreenstein).gg(1.0))


def test_workload_flop_and_byte_model():
    # BASELINE.md section 3 table
    for name, kw, w, q in (("amars_sw", dict(nwave=2, ncol=1, nlyr=40), 296053, 1324),
                           ("amars_lw", dict(nwave=2, ncol=1, nlyr=80), 602187, 2604),
                           ("amars_lw", dict(nwave=2, ncol=1, nlyr=40), 301173, None),
                           ("lbl", dict(nwave=2, ncol=1, nlyr=40), 37595, 1324),
                           ("gcm", dict(nwave=1, ncol=2, nlyr=60), 3530176, 9644)):
        wl = W.make(name, **kw)
        assert abs(wl.flops_per_solve() - w) < 1.0
        if q:
            assert wl.bytes_per_solve() == q


def test_workloads_are_deterministic_and_valid():
    for name, kw in (("amars_sw", dict(nwave=20, ncol=3)), ("amars_lw", dict(nwave=8, ncol=3)),
                     ("lbl", dict(nwave=100)), ("gcm", dict(nwave=2, ncol=16))):
        a, b = W.make(name, **kw), W.make(name, **kw)
        assert np.array_equal(a.prop, b.prop)
        assert (a.prop[..., 0] >= 0).all()
        if a.prop.shape[3] > 1:
            assert ((a.prop[..., 1] >= 0) & (a.prop[..., 1] <= 1)).all()
        assert not np.array_equal(a.prop, W.make(name, variant=1, **kw).prop)
