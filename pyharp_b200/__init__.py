"""pyharp_b200: B200-native DISORT solve path behind pyharp's solver plugin API."""
from .disort import (Disort, DisortOptions, DisortState, PhaseMomentOptions, scattering_moments,
                     fp64_peak_tflops, IEX, ISS, IPM, IUP, IDN, kIsotropic, kRayleigh,
                     kHenyeyGreenstein)

from .band import RadiationBand, RadiationBandOptions  # noqa: E402

__all__ = ["RadiationBand", "RadiationBandOptions", "Disort", "DisortOptions", "DisortState", "PhaseMomentOptions", "scattering_moments",
           "fp64_peak_tflops", "IEX", "ISS", "IPM", "IUP", "IDN", "kIsotropic", "kRayleigh",
           "kHenyeyGreenstein"]
