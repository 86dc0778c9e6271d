"""B200-native ELAS stereo hot path: CUDA kernels + C-ABI in csrc/, ctypes host mirror in binding.py.

The directory name is fixed by the build contract and is not a Python identifier; load it by path
(tests/conftest.py: load_binding()) under the module name `elas_b200`.
"""
from . import binding, sharding, sv  # noqa: F401  (sv: the reference's `stereo_vision` plugin class on this library)
