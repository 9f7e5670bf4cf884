"""emagls_b200 -- B200-native engine for the eMagLS filter-design path and binaural render.

The package holds the CUDA sources (csrc/), the C-ABI shared library they build into
(lib/libemagls_cuda.so) and the host-side mirror of the reference's MATLAB interface (api.py).
"""
from .api import *  # noqa: F401,F403
