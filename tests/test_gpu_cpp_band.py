"""C++ drop-in evidence for the band-level path, on the GPU (binaries prebuilt by tests/cpp/Makefile):

  * the REFERENCE's examples/amars_lw.cpp compiled UNCHANGED -- its RFM table reader (src/opacity/rfm.cpp:34-120) and
    read_weights_rfm (src/utils/read_weights.cpp:18-46) run their NetCDF branches through the stand-in
    tests/cpp/shim_nc/netcdf.h over a synthetic amarsw-ck-B1.nc, its disort::Disort is this repository's;
  * harp::RadiationBand as the reference's header declares it, with its members defined by the drop-in translation
    unit pyharp_b200/csrc/host/radiation_band_b200.cpp (one b200_band_forward call per forward()).

Both are checked against the chain of reference pieces (tests/band_chain.py)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import band_chain as BC
import nc_container as NC
import oracle_lib as O
import ref_harp_lib as R
from pyharp_b200 import workloads as W

pytestmark = pytest.mark.gpu
BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "build")
RUN = os.path.join(BUILD, "run")


def _exe(name):
    path = os.path.join(BUILD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} was not built (no /root/reference at build time)")
    return path


def _as_reference_reads_it(bw):
    """The pressure axis makes a round trip through the file (Pressure in Pa, RFMImpl::reset takes its log with
    torch): the chain uses the axis exactly as the reference ends up holding it."""
    nw = bw.nwave
    for t in bw.tables:
        npres = t["kdata"].shape[1]
        lnp = torch.from_numpy(np.exp(t["kaxis"][nw:nw + npres])).log().numpy()
        t["kaxis"] = np.concatenate([t["kaxis"][:nw], lnp, t["kaxis"][nw + npres:]])
        t["krefatm"] = np.stack([lnp, t["krefatm"][1]])
    return bw


def _records(path):
    recs, cur = [], None
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            tok = line.decode().split()
            if tok[0] == "forward":
                cur = {"flags": tok[1], "nstr": int(tok[3])}
                recs.append(cur)
                continue
            shape = [int(x) for x in tok[2:2 + int(tok[1])]]
            cur[tok[0]] = np.frombuffer(f.read(8 * int(np.prod(shape))), dtype=np.float64).reshape(shape).copy()
    return recs


def test_reference_amars_lw_unchanged(cuda_device):
    exe = _exe("ref_amars_lw")
    bw = W.make_band("amars_lw", nwave=16, ncol=1, nlyr=40, variant=3)
    os.makedirs(RUN, exist_ok=True)
    NC.write_ck_file(os.path.join(RUN, "amarsw-ck-B1.nc"), bw)
    bw = _as_reference_reads_it(bw)
    dump = os.path.join(RUN, "amars_lw_dump.bin")
    if os.path.exists(dump):
        os.remove(dump)
    r = subprocess.run([exe], cwd=RUN, capture_output=True, text=True, timeout=600, env=dict(os.environ, B200DISORT_DUMP=dump))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "prop shape = [16, 1, 40, 1]" in r.stdout and "bflx =" in r.stdout and "nan" not in r.stdout.lower()
    rec = _records(dump)[0]
    # the example's inputs (examples/amars_lw.cpp:46-76): conc = 1, p = 10e5 Pa, T = 300 K, albedo 1, btemp 300
    conc, pres, temp = np.ones((1, 40, 2)), np.full((1, 40), 10.e5), np.full((1, 40), 300.0)
    prop = sum(R.rfm_forward(t["kaxis"], t["kdata"], t["krefatm"], pres, temp, conc, t["species_id"]) for t in bw.tables)
    np.testing.assert_allclose(rec["prop"], prop, rtol=1e-13, atol=0)  # (the reference's own CPU code on both sides)
    np.testing.assert_array_equal(rec["temf"], R.layer2level(temp))
    ref, _, st = O.forward(rec["prop"], {"albedo": rec["bc_albedo"], "btemp": rec["bc_btemp"]}, rec["temf"], nstr=8,
                           flags=O.LAMBER | O.PLANCK | O.ONLYFL, wave_lower=np.full(16, 1.0), wave_upper=np.full(16, 150.0),
                           long_double=True)
    assert not st.any()
    scale = np.max(np.abs(ref), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(rec["flx"] - ref) / scale)) < 1e-9
    # the printed band flux (4 digits): sum_g w_g F_g, examples/amars_lw.cpp:84-88
    want = (ref * bw.weights[:, None, None, None]).sum(0)
    nums = [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]?\d+)?", r.stdout.split("bflx =")[1])]
    got = np.array(nums[:want.size]).reshape(want.shape)
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-4 * np.abs(want).max())


@pytest.mark.parametrize("where", ["cpu", "cuda"])
def test_radiation_band_dropin_translation_unit(cuda_device, where):
    exe = _exe("band_dropin")
    bw = W.make_band("amars_lw", nwave=16, ncol=9, nlyr=40, variant=2)
    os.makedirs(RUN, exist_ok=True)
    NC.write_ck_file(os.path.join(RUN, "amarsw-ck-B1.nc"), bw)
    bcf = bw.bc_full()
    NC.write(os.path.join(RUN, "band_inputs.nc"), {"nwave": bw.nwave, "ncol": bw.ncol, "nlyr": bw.nlyr},
             {"conc": bw.conc, "pres": bw.pres, "temp": bw.temp, "dz": bw.dz, "albedo": bcf["albedo"], "btemp": bcf["btemp"]})
    bw = _as_reference_reads_it(bw)
    r = subprocess.run([exe, where], cwd=RUN, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "BAND_DROPIN_OK" in r.stdout and "b200 band-level back end" in r.stdout
    assert f"on {'cuda:0' if where == 'cuda' else 'cpu'}" in r.stdout
    load = lambda n, shape: np.fromfile(os.path.join(RUN, n), dtype=np.float64).reshape(shape)
    flx = load("band_flx.bin", (bw.nwave, bw.ncol, bw.nlyr + 1, 2))
    prop = load("band_prop.bin", (bw.nwave, bw.ncol, bw.nlyr, 1))
    bflx = load("band_bflx.bin", (bw.ncol, bw.nlyr + 1, 2))
    prop_ref, _, flx_ref, bflx_ref = BC.lw_chain(bw)
    np.testing.assert_allclose(prop, prop_ref, rtol=1e-12, atol=0)
    scale = np.max(np.abs(flx_ref), axis=(2, 3), keepdims=True)
    assert float(np.max(np.abs(flx - flx_ref) / scale)) < 1e-9
    assert float(np.max(np.abs(bflx - bflx_ref)) / np.max(np.abs(bflx_ref))) < 1e-9
