"""``Models.fno.spectral_layers`` / ``fno.spectral_layers`` -> deno4pytorch_b200.spectral_layers."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from deno4pytorch_b200.spectral_layers import *  # noqa: F401,F403,E402
from deno4pytorch_b200.spectral_layers import (SpectralConv1d, SpectralConv2d, SpectralConv3d, AdaptiveFourier1d, AdaptiveFourier2d,
                                               AdaptiveFourier3d, F, Parameter,  # noqa
                                               activation_dict, default, fft, math, nn, partial, torch)
