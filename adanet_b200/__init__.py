"""adanet_b200: a B200-native AdaNet candidate-training engine behind the
tensorflow/adanet API surface (adanet/__init__.py:21-59 of the reference).

The per-iteration hot path runs as hand-written sm_100a CUDA kernels
(adanet_b200/csrc, C ABI in include/adanet_b200.h); this package is the
host-side mirror of the reference's plugin interface over it.
"""

from adanet_b200 import distributed

__version__ = "0.1.0"
