"""The band-level forward (include/b200band.h, pyharp_b200.RadiationBand): opacity -> mixing -> layer2level + Planck ->
DISORT solve -> fused band integration as ONE call, against the chain of its reference pieces (tests/band_chain.py:
the reference's own RFM / layer2level code from oracle/_ref + the oracle's DISORT + the weighted sum of
examples/amars_lw.cpp:88).  Tolerances: optical properties rtol 1e-14 (exact interpolation; exp differs in the last
ulp) / 1e-12 (hoisted blend weights); fluxes 1e-9 of the largest flux of the solve (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

import band_chain as BC
import golden_lib as G
import ref_harp_lib as R
from pyharp_b200 import Disort, DisortOptions, RadiationBand, RadiationBandOptions, opacity as OP, workloads as W

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _rel(a, ref):
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


@pytest.fixture(scope="module")
def lw():
    if not R.available():
        pytest.fail("oracle/_ref/libharp_ref.so missing (build() makes it from /root/reference)")
    bw = W.make_band("amars_lw", nwave=16, ncol=37, nlyr=40, variant=1)
    return bw, BC.lw_chain(bw)


@pytest.mark.parametrize("exact", [True, False], ids=["exact_interp", "hoisted_weights"])
def test_band_lw_against_reference_chain(cuda_device, lw, exact):
    bw, (prop_ref, temf_ref, flx_ref, bflx_ref) = lw
    band = W.band_of(bw, exact_interp=exact)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    op = {"pres": t(bw.pres), "temp": t(bw.temp)}
    bc = {k: t(v) for k, v in bw.bc.items()}
    # the reference's own return value: per-spectral-point fluxes, optics exported
    flx = band.forward(t(bw.conc), t(bw.dz), bc, op).cpu().numpy()
    prop = band.shared["radiation/B1/optics"].cpu().numpy()
    np.testing.assert_allclose(prop, prop_ref, rtol=1e-14 if exact else 1e-12, atol=0)
    scale = np.max(np.abs(flx_ref), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(flx - flx_ref) / scale)) < RTOL
    assert _rel(band.bflx.cpu().numpy(), bflx_ref) < RTOL
    # fused band integration: device entry, then host entry (chunked), full-shape boundary fields
    bflx = band.forward_band(t(bw.conc), t(bw.dz), bc, op).cpu().numpy()
    assert _rel(bflx, bflx_ref) < RTOL
    assert _rel(bflx, (flx * bw.weights[:, None, None, None]).sum(0)) < 1e-13
    f = lambda a: torch.from_numpy(a)
    bflx_h = band.forward_band(f(bw.conc), f(bw.dz), {k: f(v) for k, v in bw.bc_full().items()},
                               {"pres": f(bw.pres), "temp": f(bw.temp)}, want_status=True)
    assert not bflx_h.is_cuda and bflx_h.is_pinned()
    assert _rel(bflx_h.numpy(), bflx) < 1e-13
    assert int(band.status().abs().sum()) == 0
    assert band.stats()["solves_timed"] >= 1


@pytest.mark.parametrize("shape", [dict(nwave=70, ncol=3, nlyr=17), dict(nwave=5, ncol=1, nlyr=33), dict(nwave=33, ncol=65, nlyr=9)],
                         ids=["70x3x17", "5x1x33", "33x65x9"])
def test_band_odd_shapes(cuda_device, shape):
    """Ragged tiles of the optics kernel (spectral points not a multiple of 32, fewer than 32 points), one column
    (warp-reduced band integration, warp-per-solve flux kernel), host and device entries."""
    bw = W.make_band("amars_lw", variant=4, **shape)
    prop_ref, _, flx_ref, bflx_ref = BC.lw_chain(bw)
    band = W.band_of(bw)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    f = torch.from_numpy
    op = {"pres": t(bw.pres), "temp": t(bw.temp)}
    flx = band.forward(t(bw.conc), t(bw.dz), {k: t(v) for k, v in bw.bc.items()}, op).cpu().numpy()
    np.testing.assert_allclose(band.shared["radiation/B1/optics"].cpu().numpy(), prop_ref, rtol=1e-12, atol=0)
    scale = np.max(np.abs(flx_ref), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(flx - flx_ref) / scale)) < RTOL
    for where in ("device", "host"):
        mv = t if where == "device" else f
        b = band.forward_band(mv(bw.conc), mv(bw.dz), {k: mv(v) for k, v in bw.bc.items()}, {"pres": mv(bw.pres), "temp": mv(bw.temp)})
        assert _rel(b.cpu().numpy(), bflx_ref) < RTOL, where


def test_band_host_chunks_and_heating_rate(cuda_device, lw):
    """Several chunks through the three slots (chunk = 8 columns of 37, last one ragged), heating rate output."""
    bw, (_, _, _, bflx_ref) = lw
    atts_band = W.band_of(bw)
    whole = atts_band.forward_band(torch.from_numpy(bw.conc).to(cuda_device), torch.from_numpy(bw.dz),
                                   {k: torch.from_numpy(v) for k, v in bw.bc.items()},
                                   {"pres": torch.from_numpy(bw.pres), "temp": torch.from_numpy(bw.temp)}).cpu().numpy()
    band = W.band_of(bw)
    band.options.host_chunk_cols(8)
    band2 = RadiationBand(band.options, nspecies=2)
    rho = bw.pres / (188.9 * bw.temp)
    f = lambda a: torch.from_numpy(a)
    for _ in range(2):  # second call reuses the slots
        bflx, dT = band2.forward_band(f(bw.conc), f(bw.dz), {k: f(v) for k, v in bw.bc.items()},
                                      {"pres": f(bw.pres), "temp": f(bw.temp)}, rho=f(rho), cp=844.0)
        assert _rel(bflx.numpy(), whole) < 1e-13
        assert _rel(bflx.numpy(), bflx_ref) < RTOL
        np.testing.assert_allclose(dT.numpy(), BC.heating_rate(bflx.numpy(), rho, bw.dz, 844.0), rtol=1e-12, atol=1e-30)


def test_band_bad_input_is_reported(cuda_device, lw):
    bw, _ = lw
    band = W.band_of(bw)
    f = lambda a: torch.from_numpy(a)
    temp = bw.temp.copy()
    temp[3, :] = -50.0  # negative interface temperature: layer2level's positivity check / c_check_inputs
    with pytest.raises(RuntimeError, match="failed"):
        band.forward_band(f(bw.conc), f(bw.dz), {k: f(v) for k, v in bw.bc.items()}, {"pres": f(bw.pres), "temp": f(temp)})
    # and the handle is usable afterwards
    out = band.forward_band(f(bw.conc), f(bw.dz), {k: f(v) for k, v in bw.bc.items()}, {"pres": f(bw.pres), "temp": f(bw.temp)})
    assert torch.isfinite(out).all()


def _aerosol_band(z, nwave, ncol, nlyr, mix_eps=0.0):
    atts = {}
    for kind, cls, sid in (("s8_fuller", OP.S8Fuller, 0), ("h2so4_simple", OP.H2SO4Simple, 1)):
        op = OP.AttenuatorOptions().species_weights(list(z["species_weights"])).species_names(["S8", "H2SO4"]).species_ids([sid])
        atts[kind] = cls(op, table=z[kind + "_table"])
    dop = DisortOptions().flags("lamber,quiet,onlyfl,intensity_correction,old_intensity_correction")
    dop.ds().nstr = dop.ds().nmom = 8
    bo = RadiationBandOptions().attenuators(atts).disort(dop).nlyr(nlyr).ncol(ncol).mix_eps(mix_eps)
    bo.wave_lower([0.0] * nwave).wave_upper([0.0] * nwave)
    return bo, atts


@pytest.mark.parametrize("ncol", [1, 4, 40], ids=["one_column_warp_reduced", "four_columns_warp_reduced", "atomics"])
def test_band_sw_aerosols(cuda_device, ncol):
    """amars_sw as a band (examples/amars_sw.cpp:218-301): S8 + H2SO4 aerosols, beam source; against the same stages
    run one by one through the stand-alone kernels (which are pinned bit-exact against the reference)."""
    z = np.load(os.path.join(G.GOLDEN_DIR, "opacity_aerosol.npz"))
    wn = z["wavenumber"]
    nwave, nlyr = len(wn), 24
    rng = np.random.default_rng(5 + ncol)
    conc = np.abs(rng.normal(1.0, 0.5, (ncol, nlyr, 2))) * 3e-3 * np.exp(-np.linspace(0, 5, nlyr))[None, :, None]
    conc[:, -3:, :] = 0.0                      # nothing aloft: extinction 0 -> ssa 0, not 0/0
    dz = np.linspace(400.0, 1600.0, nlyr)
    weights = rng.uniform(0.5, 1.5, nwave)
    bo, atts = _aerosol_band(z, nwave, ncol, nlyr)
    bo.weights(weights)
    band = RadiationBand(bo, wave=wn, wave_is_wavenumber=True)
    assert band.nprop == 2
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    bc = {"fbeam": t(rng.uniform(0.5, 2.0, (nwave, ncol))), "umu0": t(np.full((nwave, ncol), 0.8)),
          "albedo": t(np.full((nwave, ncol), 0.3))}
    flx = band.forward(t(conc), t(dz), bc, {})
    prop = band.shared["radiation/B1/optics"]
    # stage by stage
    outs = [m.forward(t(conc), {"wavenumber": t(wn)}) for m in atts.values()]
    prop_ref = OP.band_optics(outs, t(dz))
    assert torch.equal(prop[..., 0], prop_ref[..., 0])
    np.testing.assert_allclose(prop[..., 1].cpu().numpy(), prop_ref[..., 1].cpu().numpy(), rtol=4e-16, atol=0)
    assert float(prop[:, :, -3:, :].abs().max()) == 0.0
    ds = Disort(band.options.disort())
    flx_ref = ds.forward(prop_ref, bc)
    assert _rel(flx.cpu().numpy(), flx_ref.cpu().numpy()) < 1e-12
    want = (flx_ref.cpu() * torch.from_numpy(weights).view(-1, 1, 1, 1)).sum(0).numpy()
    got = band.forward_band(t(conc), t(dz), bc, {}).cpu().numpy()
    assert _rel(got, want) < 1e-12
    got_h = band.forward_band(torch.from_numpy(conc), torch.from_numpy(dz), {k: v.cpu() for k, v in bc.items()}, {}).numpy()
    assert _rel(got_h, want) < 1e-12


def test_radiation_band_guarded_mixing(cuda_device):
    """mix_eps = 1e-10: ssa = sum(ssa ext) / (sum(ext) + 1e-10), tau = sum(ext) dz (radiation_band.cpp:107-116)."""
    z = np.load(os.path.join(G.GOLDEN_DIR, "opacity_aerosol.npz"))
    wn = z["wavenumber"]
    nwave, ncol, nlyr = len(wn), 3, 6
    conc = np.full((ncol, nlyr, 2), 2e-3)
    dz = np.full(nlyr, 700.0)
    bo, atts = _aerosol_band(z, nwave, ncol, nlyr, mix_eps=1e-10)
    band = RadiationBand(bo, wave=wn)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    bc = {"fbeam": t(np.ones((nwave, ncol))), "umu0": t(np.full((nwave, ncol), 0.5))}
    band.forward(t(conc), t(dz), bc, {})
    prop = band.shared["radiation/B1/optics"].cpu().numpy()
    outs = [m.forward(t(conc), {"wavenumber": t(wn)}).cpu().numpy() for m in atts.values()]
    ext, sse = outs[0][..., 0] + outs[1][..., 0], outs[0][..., 1] + outs[1][..., 1]
    np.testing.assert_array_equal(prop[..., 0], ext * dz[None, None, :])
    np.testing.assert_array_equal(prop[..., 1], np.minimum(sse / (ext + 1e-10), 1.0))


@pytest.mark.parametrize("name,kw", [("amars_lw", dict(nwave=32, ncol=70, nlyr=30)), ("lbl", dict(nwave=700, ncol=1, nlyr=20)),
                                     ("lbl", dict(nwave=300, ncol=8, nlyr=20)), ("gcm", dict(nwave=4, ncol=40, nlyr=12)),
                                     ("amars_sw", dict(nwave=40, ncol=3, nlyr=20)),
                                     # batches in the layer-split form's range at nstr = 8 (team adding kernel): one column =
                                     # replicated band-flux buffer, 40 columns = plain reductions
                                     ("amars_sw", dict(nwave=3001, ncol=1, nlyr=20)), ("amars_lw", dict(nwave=64, ncol=40, nlyr=24))],
                         ids=["lw_nstr8", "lbl_nstr4_one_column", "lbl_nstr4_eight_columns", "gcm_nstr16", "sw_three_columns",
                              "sw_split_team_one_column", "lw_split_team_forty_columns"])
def test_solve_band_entry(cuda_device, name, kw):
    """b200disort_solve_band: solve + fused band integration == reduce_band(forward()) (the two-pass form)."""
    wl = W.make(name, **kw)
    ds = Disort(wl.options())
    t = lambda a: None if a is None else torch.from_numpy(a).to(cuda_device)
    prop, bc, temf = t(wl.prop), {k: t(v) for k, v in wl.bc.items()}, t(wl.temf)
    w = torch.rand(wl.shape[0], dtype=torch.float64) + 0.1
    want = ds.reduce_band(ds.forward(prop, bc, temf), w).cpu().numpy()
    got = ds.forward_band(prop, bc, temf, w).cpu().numpy()
    assert _rel(got, want) < 1e-13
