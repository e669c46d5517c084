"""B200-native implementation of the C2-Ray3Dm photo-ionization hot path (evolve3D).

The compute lives in libc2ray_b200.so (hand-written CUDA for sm_100a behind the C ABI of
include/c2ray_b200.h); this package is the thin host side that mirrors the reference's
module-state interface.  Importing the package does not load the library; constructing
`Evolve` does and fails loudly when it is missing.
"""
from . import constants  # noqa: F401
from .evolve import Evolve, C2RayError, shard_sources  # noqa: F401
