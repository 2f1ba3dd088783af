"""distributed_sddmm_b200 -- B200-native SDDMM / SpMM engine behind the Half-and-Half
(PASSIONLab/distributed_sddmm) plugin surface.

The product is the CUDA / C++ shared library ``libhnh_b200.so`` built in-tree from
``csrc/`` (see ``build.py``); this Python package is only the loader and a thin ctypes
binding used by bench.py and the tests.  There is NO CPU fallback: if the library has not
been built, ``lib()`` raises.
"""
from ._lib import lib, library_path, LibraryMissing, check  # noqa: F401

__all__ = ["lib", "library_path", "LibraryMissing", "check"]
