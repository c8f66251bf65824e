"""General path (K3 device code compiled for the host) against the published DISORT test-1a
intensities and against the oracle on multi-layer, anisotropic, thermal and azimuth-dependent cases."""
import json
import os

import numpy as np
import pytest

import golden_lib as G
import hostsim_lib as H
import intensity_cases as IC
import oracle_lib as O


def test_test1a_intensities():
    t = json.load(open(os.path.join(G.GOLDEN_DIR, "test1a.json")))
    nstr = t["nstr"]
    prop = np.zeros((10, 1, 1, 2 + nstr))
    prop[..., 0], prop[..., 1] = t["dtau"], t["ssalb"]
    bc = {"umu0": np.full((10, 1), t["umu0"]), "fbeam": np.full((10, 1), np.pi * t["fbeam_over_pi"])}
    flags = O.USRTAU | O.USRANG | O.LAMBER | O.INTCOR | O.OLDINTCOR
    flx, uu, st = H.intensity(prop, bc, nstr=nstr, flags=flags, utau=t["utau"], umu=t["umu"], phi=t["phi"])
    assert not st.any()
    for n in range(10):
        np.testing.assert_allclose(uu[n, 0, 0], np.array(t["uu"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(flx[n, 0, ::-1, 0], t["flup"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(flx[n, 0, ::-1, 1], np.array(t["rfldir"]) + np.array(t["rfldn"]), rtol=1e-9)


@pytest.mark.parametrize("case", IC.CASES, ids=lambda c: "-".join(map(str, c)))
def test_general_path_vs_oracle(case):
    kw = IC.make(case)
    a = O.forward(**kw)
    b = H.intensity(**kw)
    assert not a[2].any() and not b[2].any()
    assert np.max(np.abs(a[0] - b[0])) / np.max(np.abs(a[0])) < 1e-9
    assert np.max(np.abs(a[1] - b[1])) / np.max(np.abs(a[1])) < 1e-9


def test_azimuth_convergence_rule():
    """accur > 0 stops the Fourier series after two consecutive converged modes, like c_disort."""
    kw = IC.make((16, 3, 16, 1, 0, 1, 0, 0))
    kw["accur"] = 1e-3
    a = O.forward(**kw)
    b = H.intensity(**kw)
    assert np.max(np.abs(a[1] - b[1])) / np.max(np.abs(a[1])) < 1e-9
    kw["accur"] = 0.0
    c = H.intensity(**kw)
    assert np.max(np.abs(c[1] - b[1])) / np.max(np.abs(c[1])) > 1e-9  # the truncated series really differs


def test_general_path_status_codes():
    kw = IC.make((8, 5, 8, 1, 0, 1, 1, 0))
    kw["prop"][0, 1, 2, 1] = 1.7
    kw["utau"] = [0.0, 1e9]  # beyond the bottom of the medium
    a = O.forward(**kw)
    b = H.intensity(**kw)
    assert (a[2] != 0).all() and (b[2] != 0).all()
    assert np.isnan(b[0]).all()
