"""GPU opacity / interpolation / layer2level kernels (include/b200opacity.h) against fixtures produced by the
REFERENCE's own code (tests/golden/make_golden_opacity.py).  Interpolation and aerosol optics are required to be
BIT-IDENTICAL (same arithmetic, no FMA contraction); RFM goes through exp() and layer2level's 4th-order stencil
through a BLAS matmul in the reference, so those use rtol = 1e-14."""
import os

import numpy as np
import pytest
import torch

import golden_lib as G
from pyharp_b200 import Disort, DisortOptions, opacity as OP

pytestmark = pytest.mark.gpu


def _load(name):
    return np.load(os.path.join(G.GOLDEN_DIR, name))


def test_interpn_bit_exact(cuda_device):
    z = _load("opacity_interpn.npz")
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_out")})
    for key in keys:
        ndim = int(key[2])
        axes = [z[f"{key}_axis{k}"] for k in range(ndim)]
        got = OP.interpn(z[key + "_coord"], z[key + "_data"], axes).cpu().numpy()
        assert np.array_equal(got, z[key + "_out"]), key


def test_aerosol_optics_bit_exact(cuda_device):
    z = _load("opacity_aerosol.npz")
    for kind, cls, sid in (("s8_fuller", OP.S8Fuller, 0), ("h2so4_simple", OP.H2SO4Simple, 1)):
        op = OP.AttenuatorOptions().species_weights(list(z["species_weights"])).species_names(["S8", "H2SO4"])
        op.species_ids([sid])
        m = cls(op, table=z[kind + "_table"])
        got = m.forward(torch.from_numpy(z["conc"]), {"wavenumber": torch.from_numpy(z["wavenumber"])})
        assert got.is_cuda and tuple(got.shape) == z[kind + "_out"].shape
        assert np.array_equal(got.cpu().numpy(), z[kind + "_out"]), kind
        got = m.forward(torch.from_numpy(z["conc"]), {"wavelength": torch.from_numpy(z[kind + "_wavelength"])})
        assert np.array_equal(got.cpu().numpy(), z[kind + "_out_wl"]), kind
    with pytest.raises(ValueError):
        m.forward(torch.from_numpy(z["conc"]), {})


def test_rfm_optics(cuda_device):
    z = _load("opacity_rfm.npz")
    for sid in (0, 1):
        op = OP.AttenuatorOptions().species_ids([sid])
        m = OP.RFM(op, z["kaxis"], z["kdata"], z["krefatm"])
        got = m.forward(torch.from_numpy(z["conc"]), {"pres": torch.from_numpy(z["pres"]),
                                                      "temp": torch.from_numpy(z["temp"])}).cpu().numpy()
        np.testing.assert_allclose(got, z[f"out{sid}"], rtol=1e-14, atol=0)


def test_layer2level(cuda_device):
    z = _load("opacity_layer2level.npz")
    for k in z.files:
        if k.startswith("out"):
            nlyr, o, bl, bu = k[3:].split("_")
            opt = OP.Layer2LevelOptions(order=int(o[1:]), blower=int(bl[1:]), bupper=int(bu[1:]))
            got = OP.layer2level(torch.from_numpy(z["var" + nlyr]).to(cuda_device), opt).cpu().numpy()
            np.testing.assert_allclose(got, z[k], rtol=1e-14, atol=0, err_msg=k)
    with pytest.raises(RuntimeError, match="check failed"):
        OP.layer2level(torch.tensor([[1.0, 10.0, 1.0]], dtype=torch.float64, device=cuda_device),
                       OP.Layer2LevelOptions(order=2, blower=OP.kExtrapolate))


def test_optics_feed_the_solver_on_device(cuda_device):
    """amars_sw pipeline (examples/amars_sw.cpp:218-280) with every stage on the GPU: aerosol optics ->
    band mixing -> DISORT, no host round-trip; checked against the same stages done with torch ops."""
    z = _load("opacity_aerosol.npz")
    conc = torch.from_numpy(z["conc"]).to(cuda_device) * 1e4
    wn = torch.from_numpy(z["wavenumber"]).to(cuda_device)
    atts = []
    for kind, cls, sid in (("s8_fuller", OP.S8Fuller, 0), ("h2so4_simple", OP.H2SO4Simple, 1)):
        op = OP.AttenuatorOptions().species_weights(list(z["species_weights"])).species_ids([sid])
        atts.append(cls(op, table=z[kind + "_table"]).forward(conc, {"wavenumber": wn}))
    nwave, ncol, nlyr, _ = atts[0].shape
    dz = torch.linspace(500.0, 1500.0, nlyr, dtype=torch.float64, device=cuda_device)
    prop = OP.band_optics(atts, dz.view(-1, 1))
    ref = atts[0] + atts[1]
    ref = ref * dz.view(1, 1, -1, 1)
    ref[..., 1] = ref[..., 1] / ref[..., 0]
    torch.testing.assert_close(prop, ref, rtol=1e-14, atol=0)
    assert prop.is_cuda
    o = DisortOptions().flags("lamber,quiet,onlyfl").nwave(nwave).ncol(ncol)
    o.ds().nlyr, o.ds().nstr, o.ds().nmom = nlyr, 8, 8
    bc = {"fbeam": torch.ones((nwave, ncol), dtype=torch.float64, device=cuda_device),
          "umu0": torch.ones((nwave, ncol), dtype=torch.float64, device=cuda_device),
          "albedo": torch.ones((nwave, ncol), dtype=torch.float64, device=cuda_device)}
    flx = Disort(o).forward(prop, bc)
    assert flx.is_cuda and bool(torch.isfinite(flx).all())
    import oracle_lib as O
    want, _, st = O.forward(prop.cpu().numpy(), {k: v.cpu().numpy() for k, v in bc.items()}, nstr=8,
                            flags=O.LAMBER | O.ONLYFL)
    assert not st.any()
    assert G.relerr(flx.cpu().numpy(), want) < 1e-9


def test_spherical_flux_correction_bit_exact(cuda_device):
    z = _load("opacity_spherical.npz")
    flx = torch.from_numpy(z["flx"]).to(cuda_device).contiguous()
    got = OP.spherical_flux_correction(flx, z["x1f"], z["area"], z["vol"]).cpu().numpy()
    assert np.array_equal(got, z["out"])


def test_heating_rate_matches_example_arithmetic(cuda_device):
    """examples/amars_sw.cpp:290-301 restated in numpy with the example's evaluation order: bit-exact.  The fluxes are
    the band integral (examples/amars_sw.cpp:173-196 = reduce_band with weight d_lambda) of the real amars_sw run."""
    z = _load("amars_sw_real.npz")
    flx = torch.from_numpy(z["flx_oracle_ld"]).to(cuda_device)           # (nwave, 1, nlyr+1, 2)
    d_lambda = 96.19238476953907
    bflx = (flx * d_lambda).sum(0)                                         # (1, nlyr+1, 2)
    nlyr = bflx.shape[-2] - 1
    rng = np.random.default_rng(3)
    rho = rng.uniform(1e-3, 1.5, size=(1, nlyr))
    dz = rng.uniform(200.0, 900.0, size=(1, nlyr))
    cp = 844.0
    got = OP.heating_rate(bflx, rho, dz, cp).cpu().numpy()
    b = bflx.cpu().numpy()
    want = np.empty((1, nlyr))
    for k in range(nlyr):
        df = b[0, k, 0] - b[0, k, 1]
        df1 = b[0, k + 1, 0] - b[0, k + 1, 1]
        want[0, k] = -(1 / (rho[0, k] * cp)) * (df1 - df) / dz[0, k]
    assert np.array_equal(got, want)
    # many columns at once
    bf = torch.from_numpy(rng.uniform(0, 500, size=(7, 5, nlyr + 1, 2))).to(cuda_device)
    r2, d2 = rng.uniform(0.01, 1.0, size=(7, 5, nlyr)), rng.uniform(100, 1000, size=(7, 5, nlyr))
    g2 = OP.heating_rate(bf, r2, d2, cp).cpu().numpy()
    bn = bf.cpu().numpy()
    net = bn[..., 0] - bn[..., 1]
    assert np.array_equal(g2, -(1 / (r2 * cp)) * (net[..., 1:] - net[..., :-1]) / d2)
