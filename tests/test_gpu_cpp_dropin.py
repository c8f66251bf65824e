"""The C++ host side on the GPU: tests/cpp/test_shim (assertions on the disort:: surface) and the
REFERENCE's own tests/test_disort.cpp and examples/amars_sw.cpp, compiled unchanged against this
repo's headers and CUDA back end by tests/cpp/Makefile (prebuilt; nothing here reads /root/reference)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "build")


def _run(exe, cwd=None):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")
    r = subprocess.run([path], cwd=cwd or BUILD, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_cpp_shim(cuda_device):
    out = _run("test_shim")
    assert "7 tests ran, 0 failed" in out


def test_reference_test_disort_unchanged(cuda_device):
    if not os.path.exists(os.path.join(BUILD, "ref_test_disort")):
        pytest.skip("reference drivers were not built (no /root/reference at build time)")
    out = _run("ref_test_disort")
    assert "[       OK ] TestDisort.scattering" in out
    nums = [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]\d+)?", out.split("rad =")[0])]
    # printed fluxes (4 significant digits): total down at the surface 2.3778, up at the top 0.0799, pi
    assert any(abs(v - 2.3778) < 2e-4 for v in nums) and any(abs(v - 3.1416) < 2e-4 for v in nums)
    rad = [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]\d+)?", out.split("rad =")[1].split("options")[0])]
    assert any(abs(v - 0.1178) < 2e-4 for v in rad) and any(abs(v - 0.1159) < 2e-4 for v in rad)


def test_reference_amars_sw_unchanged(cuda_device):
    if not os.path.exists(os.path.join(BUILD, "ref_amars_sw")):
        pytest.skip("reference drivers were not built (no /root/reference at build time)")
    run = os.path.join(BUILD, "run")
    out = _run("ref_amars_sw", cwd=run)
    dt = np.loadtxt(os.path.join(run, "dT_ds.txt"), comments="#")
    assert np.isfinite(dt).all() and dt.shape[0] >= 39
    assert "nan" not in out.lower()


def test_reference_radiation_band_holds_our_solver(cuda_device):
    """harp::Radiation / RadiationBand / S8Fuller compiled unchanged from the reference, with this repo's
    disort::Disort as the band's rtsolver AnyModule (tests/cpp/ref_radiation_band.cpp)."""
    if not os.path.exists(os.path.join(BUILD, "ref_radiation_band")):
        pytest.skip("reference drivers were not built (no /root/reference at build time)")
    out = _run("ref_radiation_band", cwd=os.path.join(BUILD, "run"))
    assert "RADIATION_BAND_OK" in out and "Solver: disort" in out
    toa = float(re.search(r"toa_down_mean = ([-0-9.e+]+)", out).group(1))
    assert abs(toa - 1.0) < 1e-12  # mu0 * fbeam at the top of the atmosphere
