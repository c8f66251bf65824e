"""Several solves per thread.  Small batches give every CUDA thread a single solve, so these tests shrink the grid
(one 32-thread CTA per SM; B200DISORT_BLOCK / B200DISORT_CTAS_PER_SM are read once per process, hence the
subprocess) until each thread walks through several solves of the grid-stride loop re-using its slice of the
back-sweep workspace, and compare every solve with the CPU oracle at 1e-9 -- including batches whose absorption
cut-off depth differs from solve to solve."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

SCRIPT = textwrap.dedent("""
    import sys
    import numpy as np, torch
    sys.path.insert(0, {here!r}); sys.path.insert(0, {root!r})
    import oracle_lib as O
    from pyharp_b200 import Disort, workloads as W

    def check(wl, tag):
        ref, _, st = O.forward(wl.prop, wl.bc, wl.temf, nstr=wl.nstr, flags=wl.flags,
                               wave_lower=wl.wave_lower, wave_upper=wl.wave_upper, nthreads=8)
        assert not st.any()
        ds = Disort(wl.options())
        t = lambda a: None if a is None else torch.from_numpy(a).cuda()
        flx = ds.forward(t(wl.prop), {{k: t(v) for k, v in wl.bc.items()}}, t(wl.temf)).cpu().numpy()
        scale = np.max(np.abs(ref), axis=(2, 3), keepdims=True)
        err = float(np.max(np.abs(flx - ref) / scale))
        assert err < 1e-9, (tag, err)
        assert int(ds.status().abs().sum()) == 0
        return err

    nthreads = 148 * 32  # upper bound of the shrunken grid
    # amars_lw shape (thermal, nstr = 8, no phase moments), amars_sw shape (beam), lbl shape (nstr = 4)
    for name, kw in (("amars_lw", dict(nwave=8, ncol=2500, nlyr=12)), ("amars_sw", dict(nwave=500, ncol=41, nlyr=10)),
                     ("lbl", dict(nwave=20011, ncol=1, nlyr=9))):
        wl = W.make(name, **kw)
        assert wl.nsolve > 4 * nthreads
        print(name, check(wl, name))
    # varying cut-off depth: scale the optical depths of every third solve so that abstau passes 10 part-way down
    wl = W.make("amars_sw", nwave=500, ncol=41, nlyr=10)
    prop = wl.prop.copy()
    flat = prop.reshape(-1, prop.shape[2], prop.shape[3])
    rng = np.random.default_rng(7)
    for k in range(0, flat.shape[0], 3):
        flat[k, :, 0] = rng.uniform(0.5, 6.0, size=flat.shape[1])  # thick layers
        flat[k, :, 1] = rng.uniform(0.0, 0.3, size=flat.shape[1])  # mostly absorbing
    wl.prop = prop
    print("mixed ncut", check(wl, "mixed ncut"))
""")


def test_several_solves_per_thread_match_oracle(cuda_device):
    env = dict(os.environ, B200DISORT_BLOCK="32", B200DISORT_CTAS_PER_SM="1")
    code = SCRIPT.format(here=HERE, root=os.path.dirname(HERE))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mixed ncut" in r.stdout
