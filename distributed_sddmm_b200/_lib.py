"""ctypes binding of include/hnh_b200.h (the C ABI of libhnh_b200.so)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LibraryMissing(RuntimeError):
    pass


def library_path() -> str:
    return os.path.join(_HERE, "libhnh_b200.so")


# name -> (restype, argtypes); every symbol declared in include/hnh_b200.h
_P = C.c_void_p
_I64 = C.c_int64
ABI = {
    "hnh_abi_version": (C.c_int, []),
    "hnh_build_info": (C.c_char_p, []),
    "hnh_last_error_string": (C.c_char_p, []),
    "hnh_launch_count": (C.c_uint64, []),
    "hnh_sddmm_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, C.c_int, C.c_int, _P]),
    "hnh_sddmm_coo_f64": (C.c_int, [_P, _P, _P, _I64, _P, _P, C.c_int, C.c_int, _P]),
    "hnh_spmm_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, C.c_int, C.c_int, _P]),
    "hnh_fused_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _P, C.c_int, C.c_int, _P]),
    "hnh_fill_f64": (C.c_int, [_P, _I64, C.c_double, _P]),
    "hnh_hadamard_f64": (C.c_int, [_P, _P, _P, _I64, _P]),
    "hnh_expand_row_idx": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "hnh_batch_dot_f64": (C.c_int, [_P, _P, _P, _I64, C.c_int, _P]),
    "hnh_row_axpy_f64": (C.c_int, [_P, _P, C.c_double, _P, _P, _I64, C.c_int, _P]),
    "hnh_vec_quotient_f64": (C.c_int, [_P, _P, C.c_double, _P, C.c_double, _I64, _P]),
    "hnh_axpby_f64": (C.c_int, [_P, C.c_double, _P, C.c_double, _P, _I64, _P]),
    "hnh_squared_norm_f64": (C.c_int, [_P, _P, _I64, _P]),
    "hnh_block_create_host": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, C.POINTER(_P)]),
    "hnh_block_destroy": (None, [_P]),
    "hnh_block_run_host": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "hnh_er_generate_host": (C.c_int64, [C.c_int, C.c_int, C.c_uint64, _I64, _I64, _P, _P, _P, _I64]),
    "hnh_coo_to_csr_host": (C.c_int, [_I64, _I64, _I64, _P, _P, _P, C.c_int, _P, _P, _P, _P]),
}


def lib():
    """Load libhnh_b200.so (built in-tree by build.py).  Raises LibraryMissing loudly when it
    has not been built -- the product has no fallback path."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise LibraryMissing(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " (nvcc, sm_100a).  distributed_sddmm_b200 has no CPU fallback.")
        L = C.CDLL(path)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int, what: str = "hnh call") -> None:
    if rc != 0:
        msg = lib().hnh_last_error_string().decode()
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")
