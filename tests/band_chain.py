"""TEST INFRASTRUCTURE: the per-band pipeline of pyharp restated from its pieces, as the checker of the fused
band-level forward (pyharp_b200.RadiationBand): the REFERENCE's own RFM / layer2level code compiled into
oracle/_ref (tests/ref_harp_lib.py), the mixing of the examples in numpy, the oracle's DISORT, and the weighted
band sum of examples/amars_lw.cpp:84-88."""
import numpy as np

import oracle_lib as O
import ref_harp_lib as R


def lw_chain(bw, nthreads=8):
    """-> (prop (nwave, ncol, nlyr, 1), temf, flx (nwave, ncol, nlyr+1, 2), bflx (ncol, nlyr+1, 2))."""
    prop = None
    for t in bw.tables:
        out = R.rfm_forward(t["kaxis"], t["kdata"], t["krefatm"], bw.pres, bw.temp, bw.conc, t["species_id"])
        prop = out if prop is None else prop + out          # examples/amars_lw.cpp:66-69
    prop = prop * bw.dz[None, None, :, None]                 # examples/amars_sw.cpp:269 (amars_lw assumes dz = 1)
    temf = R.layer2level(bw.temp)                            # examples/amars_lw.cpp:76
    flags = 0
    for tok, bit in (("lamber", 4), ("planck", 8), ("onlyfl", 16)):
        if tok in bw.flag_tokens.split(","):
            flags |= bit
    # the 80-bit build of the oracle: the upper layers of the weak g-points are optically very thin (dtau ~ 1e-9) under
    # kelvins of temperature contrast, where the double-precision reference arithmetic is itself ~1e-9 from the
    # extended-precision result (DISORT's linear-in-tau Planck source divides by dtau)
    flx, _, st = O.forward(prop, bw.bc_full(), temf, nstr=bw.nstr, flags=flags, wave_lower=bw.wave_lower,
                           wave_upper=bw.wave_upper, nthreads=nthreads, long_double=True)
    assert not st.any()
    bflx = (flx * bw.weights[:, None, None, None]).sum(0)    # examples/amars_lw.cpp:88
    return prop, temf, flx, bflx


def heating_rate(bflx, rho, dz, cp):
    """examples/amars_sw.cpp:290-301."""
    net = bflx[..., 0] - bflx[..., 1]
    return -(1.0 / (rho * cp)) * (net[:, 1:] - net[:, :-1]) / dz[None, :]
