"""B200-native transform + quantisation front end of the Rududu Image Codec (RIC).

The product is `librududu_b200.so` (C ABI in include/ric_b200.h, CUDA kernels in csrc/); this
package is the thin Python plumbing around it used by the tests and bench.py.
"""
from . import capi  # noqa: F401
from .capi import CDF53, CDF97, Context, RicError  # noqa: F401
from .synth import synth_image  # noqa: F401
