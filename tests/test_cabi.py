"""The C-ABI library loads and exports every symbol include/b200disort.h declares; without a GPU
every computing entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pyharp_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="b200disort.h", prefix="b200disort_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = _cabi.lib()
    names = declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200disort.h but not exported"
    assert sorted(_cabi.EXPORTS) == names
    onames = declared_functions("b200opacity.h", "b200_")
    assert len(onames) == 7 and sorted(_cabi.OPACITY_EXPORTS) == onames
    for n in onames:
        assert hasattr(lib, n), f"{n} declared in include/b200opacity.h but not exported"
    bnames = sorted(declared_functions("b200band.h", "b200_band_") + declared_functions("b200band.h", "b200disort_"))
    assert len(bnames) == 11 and sorted(_cabi.BAND_EXPORTS) == bnames
    for n in bnames:
        assert hasattr(lib, n), f"{n} declared in include/b200band.h but not exported"


def test_flag_values_match_header():
    src = open(os.path.join(ROOT, "include", "b200disort.h")).read()
    for name in ("USRTAU", "USRANG", "LAMBER", "PLANCK", "ONLYFL", "INTENSITY_CORRECTION",
                 "OLD_INTENSITY_CORRECTION", "QUIET"):
        m = re.search(rf"#define B200DISORT_{name} (\d+)u", src)
        assert m and int(m.group(1)) == getattr(_cabi, name)


def test_struct_layouts():
    assert C.sizeof(_cabi.BC) == 8 * C.sizeof(C.c_void_p)
    assert _cabi.Config.device.offset > _cabi.Config.wave_upper.offset
    assert _cabi.lib().b200disort_version().startswith(b"b200disort")
    assert b"no CPU fallback" in _cabi.lib().b200disort_strerror(_cabi.ERR_CUDA)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path cannot be exercised")
    cfg = _cabi.Config()
    cfg.nwave = cfg.ncol = cfg.nlyr = 1
    cfg.nstr = cfg.nmom = 4
    cfg.flags = _cabi.LAMBER | _cabi.ONLYFL
    h = C.c_void_p()
    assert _cabi.lib().b200disort_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_CUDA
    assert not h.value
    v = C.c_double()
    assert _cabi.lib().b200disort_fp64_peak(0, C.byref(v)) == _cabi.ERR_CUDA
    from pyharp_b200 import Disort, DisortOptions
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Disort(DisortOptions().flags("lamber,onlyfl"))


def test_band_structs_match_header():
    """ctypes mirrors of b200_attenuator / b200_band_config / b200_band_io: field order and sizes as a C compiler
    lays them out (compiled here with gcc from the header itself)."""
    import subprocess
    import tempfile
    src = r"""
#include <stdio.h>
#include <stddef.h>
#include "b200band.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200_attenuator), sizeof(b200_band_config), sizeof(b200_band_io),
         offsetof(b200_attenuator, kwave), offsetof(b200_band_config, mix_eps), offsetof(b200_band_io, bflx),
         offsetof(b200_band_io, prop), offsetof(b200_band_config, weights));
  return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(_cabi.Attenuator), C.sizeof(_cabi.BandConfig), C.sizeof(_cabi.BandIO), _cabi.Attenuator.kwave.offset,
            _cabi.BandConfig.mix_eps.offset, _cabi.BandIO.bflx.offset, _cabi.BandIO.prop.offset, _cabi.BandConfig.weights.offset]
    assert got == want


def test_band_create_rejects_bad_config():
    lib = _cabi.lib()
    h = C.c_void_p()
    cfg = _cabi.BandConfig()
    assert lib.b200_band_create(None, C.byref(h)) == _cabi.ERR_INVALID
    assert lib.b200_band_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID      # no attenuators
    assert lib.b200_band_forward(None, None) == _cabi.ERR_INVALID
    assert lib.b200_band_allreduce(None, None, 0, None) == _cabi.ERR_INVALID


def test_create_rejects_bad_config():
    cfg = _cabi.Config()
    h = C.c_void_p()
    lib = _cabi.lib()
    assert lib.b200disort_create(None, C.byref(h)) == _cabi.ERR_INVALID
    cfg.nwave = cfg.ncol = cfg.nlyr = 1
    cfg.nstr, cfg.nmom, cfg.flags = 5, 8, _cabi.LAMBER | _cabi.ONLYFL   # odd stream count
    assert lib.b200disort_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID
    cfg.nstr, cfg.nmom = 8, 4                                           # nmom < nstr
    assert lib.b200disort_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_INVALID
    cfg.nstr, cfg.nmom, cfg.flags = 8, 8, _cabi.ONLYFL                  # non-Lambertian surface
    assert lib.b200disort_create(C.byref(cfg), C.byref(h)) == _cabi.ERR_UNSUPPORTED
