"""mc-cnn_b200 -- B200 (sm_100a) drop-in for the stereo-method hot path of jzbontar/mc-cnn.

Contents: ``csrc/`` (hand-written CUDA kernels + the C ABI of libadcensus_b200.so,
declared in ../include/adcensus_b200.h), ``adcensus`` (host-side mirror of the
reference's Lua table of the same name) and ``pipeline`` (stereo_predict).
The directory name carries a hyphen; import it through the ``mccnn_b200`` module at
the repository root.
"""
from . import adcensus, pipeline  # noqa: F401
from .build import build  # noqa: F401
