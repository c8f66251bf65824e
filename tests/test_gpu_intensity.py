"""GPU parity of the general path (K3) through the public API, written the way
tests/test_disort.cpp drives pydisort: options -> Disort -> forward -> get_rad."""
import json
import os

import numpy as np
import pytest
import torch

import golden_lib as G
import intensity_cases as IC
import oracle_lib as O
from pyharp_b200 import Disort, DisortOptions, PhaseMomentOptions, scattering_moments, IEX, ISS, IPM, kIsotropic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("where", ["host", "device"])
def test_disort_cpp_scattering(cuda_device, where):
    """tests/test_disort.cpp:12-56 (TEST(TestDisort, scattering)) with the assertions it lacks."""
    op = DisortOptions()
    op.header("running disort example")
    op.flags("usrtau,usrang,lamber,quiet,intensity_correction,old_intensity_correction,print-input,print-phase-function")
    op.nwave(10)
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = 1, 16, 16
    op.user_mu([-1, -0.5, -0.1, 0.1, 0.5, 1]).user_phi([0]).user_tau([0, 0.03125])
    disort = Disort(op)
    ds = disort.ds()
    dev = "cpu" if where == "host" else cuda_device
    prop = torch.zeros((op.nwave(), op.ncol(), ds.nlyr, 2 + ds.nstr), dtype=torch.float64)
    prop[..., IEX] = ds.utau[1]
    prop[..., ISS] = 0.2
    prop[..., IPM:] = scattering_moments(ds.nstr, PhaseMomentOptions().type(kIsotropic))
    bc = {"umu0": torch.full((op.nwave(), op.ncol()), 0.1, dtype=torch.float64)}
    bc["fbeam"] = np.pi / bc["umu0"]
    result = disort.forward(prop.to(dev), {k: v.to(dev) for k, v in bc.items()})
    rad = disort.get_rad()
    t = json.load(open(os.path.join(G.GOLDEN_DIR, "test1a.json")))
    assert tuple(result.shape) == (10, 1, 2, 2) and tuple(rad.shape) == (10, 1, 1, 2, 6)
    result, rad = result.cpu().numpy(), rad.cpu().numpy()
    for n in range(10):
        np.testing.assert_allclose(rad[n, 0, 0], np.array(t["uu"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(result[n, 0, ::-1, 0], t["flup"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(result[n, 0, ::-1, 1], np.array(t["rfldir"]) + np.array(t["rfldn"]), rtol=1e-9)
    assert "nstr=16" in str(disort.options)


@pytest.mark.parametrize("case", IC.CASES, ids=lambda c: "-".join(map(str, c)))
def test_general_path_vs_oracle(cuda_device, case):
    kw = IC.make(case)
    ref_flx, ref_uu, st = O.forward(**kw)
    assert not st.any()
    nw, nc, nl, _ = kw["prop"].shape
    op = DisortOptions().flags(IC.flag_tokens(kw["flags"])).nwave(nw).ncol(nc)
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = nl, kw["nstr"], kw["nmom"]
    op.user_phi(kw["phi"])
    if "umu" in kw:
        op.user_mu(kw["umu"])
    if "utau" in kw:
        op.user_tau(kw["utau"])
    if "wave_lower" in kw:
        op.wave_lower(list(kw["wave_lower"])).wave_upper(list(kw["wave_upper"]))
    ds = Disort(op)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    flx = ds.forward(t(kw["prop"]), {k: t(v) for k, v in kw["bc"].items()}, t(kw["temf"]))
    uu = ds.get_rad()
    assert flx.shape == ref_flx.shape and uu.shape == ref_uu.shape
    assert np.max(np.abs(flx.cpu().numpy() - ref_flx)) / np.max(np.abs(ref_flx)) < 1e-9
    assert np.max(np.abs(uu.cpu().numpy() - ref_uu)) / np.max(np.abs(ref_uu)) < 1e-9


def test_user_levels_flux_only(cuda_device):
    """usrtau + onlyfl: fluxes at user optical depths, no intensities."""
    kw = IC.make((8, 6, 8, 1, 1, 1, 1, 0))
    kw["flags"] |= O.ONLYFL
    ref_flx, _, st = O.forward(**kw)
    nw, nc, nl, _ = kw["prop"].shape
    op = DisortOptions().flags(IC.flag_tokens(kw["flags"])).nwave(nw).ncol(nc)
    op.ds().nlyr, op.ds().nstr, op.ds().nmom = nl, kw["nstr"], kw["nmom"]
    op.user_tau(kw["utau"]).wave_lower(list(kw["wave_lower"])).wave_upper(list(kw["wave_upper"]))
    ds = Disort(op)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    flx = ds.forward(t(kw["prop"]), {k: t(v) for k, v in kw["bc"].items()}, t(kw["temf"])).numpy()
    assert flx.shape == ref_flx.shape
    assert np.max(np.abs(flx - ref_flx)) / np.max(np.abs(ref_flx)) < 1e-9
    with pytest.raises(RuntimeError):
        ds.get_rad()
