"""Host-side Python view of the B200-native NL-PartSol explicit hot path.

The product is libnlps_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/nlps_b200.h); this package only marshals numpy arrays into that ABI for
tests and bench.py.  There is no CPU fallback here.
"""
from .problem import Problem  # noqa: F401
