"""The opacity fixtures are outputs of the REFERENCE's own code: when oracle/_ref/libharp_ref.so is present
(built from /root/reference), re-run it and require bit-identical results.  Also checks the Python table
reader against the tables the reference parsed.  CPU only."""
import os

import numpy as np
import pytest

import golden_lib as G
import ref_harp_lib as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libharp_ref.so not built (no /root/reference)")


def _load(name):
    return np.load(os.path.join(G.GOLDEN_DIR, name))


def test_interpn_fixture_is_reference_output():
    z = _load("opacity_interpn.npz")
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files if k.endswith("_out")})
    assert len(keys) == 6
    for key in keys:
        ndim = int(key[2])
        axes = [z[f"{key}_axis{k}"] for k in range(ndim)]
        assert np.array_equal(R.interpn(z[key + "_coord"], z[key + "_data"], axes), z[key + "_out"])


def test_layer2level_fixture_is_reference_output():
    z = _load("opacity_layer2level.npz")
    for k in z.files:
        if k.startswith("out"):
            nlyr, o, bl, bu = k[3:].split("_")
            got = R.layer2level(z["var" + nlyr], int(o[1:]), int(bl[1:]), int(bu[1:]))
            np.testing.assert_allclose(got, z[k], rtol=1e-15)


def test_rfm_fixture_is_reference_output():
    z = _load("opacity_rfm.npz")
    for sid in (0, 1):
        got = R.rfm_forward(z["kaxis"], z["kdata"], z["krefatm"], z["pres"], z["temp"], z["conc"], sid)
        assert np.array_equal(got, z[f"out{sid}"])


def test_aerosol_tables_parsed_like_the_reference():
    z = _load("opacity_aerosol.npz")
    for kind, sid in (("s8_fuller", 0), ("h2so4_simple", 1)):
        tab = z[kind + "_table"]
        assert np.array_equal(tab[:, 0], z[kind + "_kwave"])
        assert np.array_equal(tab[:, 1] * z["species_weights"][sid], z[kind + "_kdata"][:, 0])
        assert np.array_equal(tab[:, 2], z[kind + "_kdata"][:, 1])


def test_spherical_fixture_is_reference_output():
    z = _load("opacity_spherical.npz")
    assert np.array_equal(R.spherical_flux_correction(z["flx"], z["x1f"], z["area"], z["vol"]), z["out"])
