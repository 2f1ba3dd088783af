"""Helpers for the -m gpu parity tests: torch is only the device-memory / stream plumbing."""
import numpy as np
import torch

from distributed_sddmm_b200 import check, lib


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def stream():
    return torch.cuda.current_stream().cuda_stream


def run_sddmm(csr, A, B, v0=None, flags=0, coo=False):
    L = lib()
    rs, ci, ri = dev(csr.rowStart), dev(csr.col_idx), dev(csr.row_idx)
    dA, dB = dev(A), dev(B)
    v = dev(np.zeros(csr.nnz) if v0 is None else v0)
    r = A.shape[1]
    if coo:
        check(L.hnh_sddmm_coo_f64(ri.data_ptr(), ci.data_ptr(), v.data_ptr(), csr.nnz,
                                  dA.data_ptr(), dB.data_ptr(), r, flags, stream()), "sddmm_coo")
    else:
        check(L.hnh_sddmm_f64(rs.data_ptr(), ci.data_ptr(), v.data_ptr(), csr.rows, csr.nnz,
                              dA.data_ptr(), dB.data_ptr(), r, flags, stream()), "sddmm")
    torch.cuda.synchronize()
    return v.cpu().numpy()


def run_spmm(csr, vals, X, Y0, flags=0):
    L = lib()
    rs, ci = dev(csr.rowStart), dev(csr.col_idx)
    dv, dX, dY = dev(vals), dev(X), dev(Y0)
    check(L.hnh_spmm_f64(rs.data_ptr(), ci.data_ptr(), dv.data_ptr(), csr.rows, csr.nnz,
                         dX.data_ptr(), dY.data_ptr(), X.shape[1], flags, stream()), "spmm")
    torch.cuda.synchronize()
    return dY.cpu().numpy()


def run_fused(csr, v0, X, Y, Out0, flags=0):
    L = lib()
    rs, ci = dev(csr.rowStart), dev(csr.col_idx)
    dv, dX, dY, dO = dev(v0), dev(X), dev(Y), dev(Out0)
    check(L.hnh_fused_f64(rs.data_ptr(), ci.data_ptr(), dv.data_ptr(), csr.rows, csr.nnz,
                          dX.data_ptr(), dY.data_ptr(), dO.data_ptr(), X.shape[1], flags, stream()),
          "fused")
    torch.cuda.synchronize()
    return dv.cpu().numpy(), dO.cpu().numpy()
