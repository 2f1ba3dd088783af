"""The C-ABI library loads on a CPU-only box and exports every symbol include/hnh_b200.h
declares; argument validation works without a GPU; compute fails loudly without one."""
import ctypes as C
import os
import re

import pytest

import distributed_sddmm_b200 as pkg
from distributed_sddmm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if hdr.endswith(".h"):
            text = open(os.path.join(ROOT, "include", hdr)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(hnhd?_[a-z0-9_]+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported(hnh):
    declared = _declared_symbols()
    assert len(declared) >= 15
    missing = [n for n in sorted(declared) if not hasattr(hnh, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    unbound = [n for n in sorted(declared) if n not in _lib.ABI]
    assert not unbound, f"declared but not bound in _lib.ABI: {unbound}"


def test_version_and_build_info(hnh):
    assert hnh.hnh_abi_version() >= 1
    assert b"sm_100a" in hnh.hnh_build_info()


def test_argument_validation_needs_no_gpu(hnh):
    assert hnh.hnh_sddmm_f64(None, None, None, 4, -1, None, None, 8, 0, None) == -1
    assert b"negative" in hnh.hnh_last_error_string()
    assert hnh.hnh_spmm_f64(None, None, None, 4, 4, None, None, 0, 0, None) == -1
    assert hnh.hnh_fused_f64(None, None, None, 4, 4, None, None, None, 8, 0, None) == -1
    # empty block: successful no-op, like sparse_kernels.cpp:25-27,85-87
    assert hnh.hnh_sddmm_f64(None, None, None, 4, 0, None, None, 8, 0, None) == 0
    assert hnh.hnh_spmm_f64(None, None, None, 0, 0, None, None, 8, 0, None) == 0


def test_round2_entry_points_validate_without_a_gpu(hnh):
    """The scaled-epilogue kernels and the device tuple pipeline reject bad arguments before touching the device."""
    assert hnh.hnh_sddmm_scaled_f64(None, None, None, 4, -1, None, None, 8, 0, None, None, None) == -1
    # scaled_out without scale
    buf = (C.c_double * 8)()
    assert hnh.hnh_sddmm_scaled_f64(None, None, None, 4, 4, None, None, 8, 0, None, buf, None) == -1
    assert b"scaled_out without scale" in hnh.hnh_last_error_string()
    assert hnh.hnh_fused_scaled_f64(None, None, None, 4, 4, None, None, None, 8, 0, None, buf, None) == -1
    # empty block: successful no-op
    assert hnh.hnh_sddmm_scaled_f64(None, None, None, 4, 0, None, None, 8, 0, None, None, None) == 0
    starts = (C.c_int64 * 3)()
    assert hnh.hnh_tuples_bucket_by_owner_device(None, None, None, -1, 0, 1, 1, None, 1, 1, 2, None, None, None, starts, None) == -1
    assert hnh.hnh_tuples_sort_colmajor_device(None, None, None, -1, 1, 1, None) == -1
    assert hnh.hnh_tuples_sort_colmajor_device(None, None, None, 1, 1, 1, None) == 0      # nothing to sort
    assert hnh.hnh_tuples_mod_device(None, None, 0, 3, 3, None) == 0
    assert hnh.hnh_tuples_block_starts_device(None, 5, 0, 2, starts, None) == -1           # zero block width
    assert hnh.hnh_tuples_block_starts_device(None, 0, 4, 2, starts, None) == 0 and list(starts) == [0, 0, 0]
    assert hnh.hnh_dgemm_f64(None, None, None, -1, 1, 1, None) == -1
    assert hnh.hnh_dgemm_f64(None, None, None, 0, 4, 4, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "library_path", lambda: str(tmp_path / "nope.so"))
    with pytest.raises(pkg.LibraryMissing):
        _lib.lib()


def test_compute_without_gpu_fails_loudly(hnh):
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    buf = (C.c_double * 64)()
    idx = (C.c_int64 * 8)()
    p = C.addressof
    rc = hnh.hnh_sddmm_f64(p(idx), p(idx), p(buf), 1, 1, p(buf), p(buf), 8, 0, None)
    assert rc == -2, "no CPU fallback: compute must fail with HNH_E_CUDA without a device"
