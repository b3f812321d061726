"""frank_b200: B200-native (sm_100a) implementation of discsim/frank's visibility -> Gaussian-process
normal-equations path, behind frank's own Python API.  See DESIGN.md."""
__version__ = "0.1.0"

from frank_b200 import constants, geometry, hankel  # noqa: F401
