"""Unit tests of two building blocks of the flux kernel (pyharp_b200/csrc/disort_device.cuh, compiled for the host by
tests/hostsim): the integer-pipe validity test of pivots / eigenvalues and the half-angle Jacobi rotation."""
import ctypes as C
import math

import numpy as np

import hostsim_lib as H


def _lib():
    L = H._lib()
    L.hostsim_pivot_bad.restype = C.c_int
    L.hostsim_pivot_bad.argtypes = [C.c_double]
    L.hostsim_jacobi_rotation.restype = None
    L.hostsim_jacobi_rotation.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)]
    return L


def test_pivot_check_accepts_exactly_the_usable_range():
    L = _lib()
    good = [1.0, 1e-289, 3.3e-290, 1e300, 1.7976931348623157e308, 2.5, 1e-100]
    bad = [0.0, -0.0, -1.0, -1e-300, 1e-291, 1e-300, 5e-324, float("inf"), float("-inf"), float("nan")]
    assert all(L.hostsim_pivot_bad(x) == 0 for x in good)
    assert all(L.hostsim_pivot_bad(x) == 1 for x in bad)
    # the threshold sits at the high word of 1e-290 (kTiny): everything from kTiny up is usable
    assert L.hostsim_pivot_bad(1.0e-290) == 0
    rng = np.random.default_rng(0)
    for x in 10.0 ** rng.uniform(-289.0, 308.0, 2000):
        assert L.hostsim_pivot_bad(float(x)) == 0
        assert L.hostsim_pivot_bad(float(-x)) == 1


def test_half_angle_rotation_annihilates_the_pivot():
    L = _lib()
    out = (C.c_double * 3)()
    rng = np.random.default_rng(1)
    eps2 = 1e-24
    for _ in range(3000):
        gpp, gqq = 10.0 ** rng.uniform(-10, 4, 2)
        gpq = rng.uniform(-1, 1) * math.sqrt(gpp * gqq) * 10.0 ** rng.uniform(-8, 0)
        L.hostsim_jacobi_rotation(gpq, gpp, gqq, eps2, out)
        t, s, c = out[0], out[1], out[2]
        if gpq * gpq <= eps2 * gpp * gqq:
            assert (t, s, c) == (0.0, 0.0, 1.0)      # below the relative threshold: the identity, not a branch
            continue
        assert abs(c * c + s * s - 1.0) < 1e-15 and abs(t - s / c) < 1e-15 * max(1.0, abs(t))
        assert c >= math.sqrt(0.5) - 1e-15 and abs(t) <= 1.0 + 1e-15     # the small-angle root
        # J^T G J with J = [[c, s], [-s, c]] zeroes the off-diagonal entry
        off = c * s * (gpp - gqq) + (c * c - s * s) * gpq
        assert abs(off) <= 1e-15 * (abs(gpp) + abs(gqq) + abs(gpq))
        # and the diagonal update used by the kernel: g_pp - t g_pq, g_qq + t g_pq
        npp = c * c * gpp - 2 * c * s * gpq + s * s * gqq
        assert abs(npp - (gpp - t * gpq)) <= 1e-14 * (abs(gpp) + abs(gqq))
    # equal diagonal entries: 45 degrees
    L.hostsim_jacobi_rotation(0.3, 2.0, 2.0, eps2, out)
    assert abs(abs(out[0]) - 1.0) < 1e-15 and abs(out[2] - math.sqrt(0.5)) < 1e-15
