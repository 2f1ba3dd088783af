"""ctypes/numpy front end of the CPU oracle (oracle/hnh_oracle.c) -- TEST INFRASTRUCTURE.

PARITY STATUS: the reference has no golden vectors for this path.  The C
restatement wrapped here is cross-checked against scipy.sparse and the dummyInitialize
closed form (tests/test_oracle.py) and, where /root/reference is present, against the
reference's own sources compiled with shims (oracle/_ref, tests/test_oracle_vs_ref.py).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "hnh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            so = build()
        L = C.CDLL(so)
        L.oracle_sddmm_coo.argtypes = [_i64p, _i64p, _f64p, C.c_int64, _f64p, _f64p, C.c_int64]
        L.oracle_sddmm_coo.restype = None
        L.oracle_spmm_csr.argtypes = [_i64p, _i64p, _f64p, C.c_int64, _f64p, _f64p, C.c_int64]
        L.oracle_spmm_csr.restype = None
        L.oracle_coo_to_csr.argtypes = [C.c_int64, C.c_int64, C.c_int64, _u64p, _u64p, _f64p,
                                        C.c_int, _i64p, _i64p, _i64p, _f64p]
        L.oracle_coo_to_csr.restype = C.c_int
        L.oracle_fused_block.argtypes = [_i64p, _i64p, _i64p, _f64p, C.c_int64, C.c_int64,
                                         _f64p, _f64p, _f64p, C.c_int64]
        L.oracle_fused_block.restype = None
        L.oracle_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ---------------------------------------------------------------- local kernels -----------
def sddmm_coo(row_idx, col_idx, values, X, Y):
    """values[i] += X[row_idx[i]] . Y[col_idx[i]]  (sparse_kernels.cpp:44-55). In place."""
    r = X.shape[1]
    assert Y.shape[1] == r  # assert(A.cols() == B.cols()), sparse_kernels.cpp:21
    lib().oracle_sddmm_coo(row_idx, col_idx, values, len(values), X, Y, r)
    return values


def spmm_csr(rowStart, col_idx, values, X, Y):
    """Y += CSR * X, alpha=beta=1, row-major (sparse_kernels.cpp:95-120). In place."""
    r = X.shape[1]
    assert Y.shape[1] == r
    lib().oracle_spmm_csr(rowStart, col_idx, values, len(rowStart) - 1, X, Y, r)
    return Y


def fused_block(rowStart, row_idx, col_idx, values, X, Y, Out):
    """K1 then K2 on the same block (15D_dense_shift.hpp:203-217). In place."""
    lib().oracle_fused_block(rowStart, row_idx, col_idx, values, len(rowStart) - 1, len(values),
                             X, Y, Out, X.shape[1])
    return values, Out


class CSR:
    """The four arrays of the reference's CSRHandle (SpmatLocal.hpp:55-62)."""

    def __init__(self, rows, cols, rowStart, col_idx, row_idx, values, transpose):
        self.rows, self.cols = rows, cols  # stored shape
        self.rowStart, self.col_idx, self.row_idx, self.values = rowStart, col_idx, row_idx, values
        self.transpose = transpose

    @property
    def nnz(self):
        return len(self.col_idx)


def coo_to_csr(rows, cols, r, c, v, transpose=False) -> CSR:
    """CSRLocal constructor (SpmatLocal.hpp:78-188): COO -> CSR, optional transpose."""
    r = np.ascontiguousarray(r, dtype=np.uint64)
    c = np.ascontiguousarray(c, dtype=np.uint64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    nnz = len(r)
    out_rows = cols if transpose else rows
    out_cols = rows if transpose else cols
    rowStart = np.zeros(out_rows + 1, dtype=np.int64)
    col_idx = np.zeros(nnz, dtype=np.int64)
    row_idx = np.zeros(nnz, dtype=np.int64)
    values = np.zeros(nnz, dtype=np.float64)
    rc = lib().oracle_coo_to_csr(rows, cols, nnz, r, c, v, int(transpose), rowStart, col_idx,
                                 row_idx, values)
    if rc != 0:
        raise ValueError(f"oracle_coo_to_csr failed ({rc}): coordinate out of range")
    return CSR(out_rows, out_cols, rowStart, col_idx, row_idx, values, transpose)


# ---------------------------------------------------------------- synthetic input ---------
_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)
_ROWMUL = np.uint64(0xD1342543DE82EF95)


def _mix(z):
    with np.errstate(over="ignore"):
        z = z + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        return z ^ (z >> np.uint64(31))


def er_tuples(logM: int, nnz_per_row: int, seed: int, row_lo: int = 0, row_hi: int | None = None):
    """Erdos-Renyi tuples of the repo's generator (numpy restatement of
    SpmatLocal::loadTuples(false, logM, nnz_per_row) in this repo, which replaces the
    CombBLAS Graph500 call at reference SpmatLocal.hpp:499-516).

    For every row i, nnz_per_row columns col = mix(mix(seed ^ i*ROWMUL) + k) mod N are drawn,
    sorted and de-duplicated.  Output is sorted by (row, col); it does not depend on how rows
    are split across ranks.  Returns (rows u64, cols u64, values f64 == 1.0)."""
    N = 1 << logM
    if row_hi is None:
        row_hi = N
    rows = np.arange(row_lo, row_hi, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = _mix(np.uint64(seed) ^ (rows * _ROWMUL))
        ks = np.arange(nnz_per_row, dtype=np.uint64)
        cols = _mix(base[:, None] + ks[None, :]) & np.uint64(N - 1)
    cols.sort(axis=1)
    keep = np.ones(cols.shape, dtype=bool)
    keep[:, 1:] = cols[:, 1:] != cols[:, :-1]
    rr = np.broadcast_to(rows[:, None], cols.shape)[keep]
    cc = cols[keep]
    return np.ascontiguousarray(rr), np.ascontiguousarray(cc), np.ones(len(rr), dtype=np.float64)


def dummy_matrix(row_lo: int, nrows: int, R: int, col_lo: int = 0, ncols: int | None = None):
    """Distributed_Sparse::dummyInitialize pattern X[g_row, g_col] = g_row*R + g_col
    (distributed_sparse.h:322-346)."""
    if ncols is None:
        ncols = R
    gr = np.arange(row_lo, row_lo + nrows, dtype=np.float64)[:, None]
    gc = np.arange(col_lo, col_lo + ncols, dtype=np.float64)[None, :]
    return np.ascontiguousarray(gr * R + gc)


def dummy_sddmm_closed_form(i, j, R):
    """sum_k (iR+k)(jR+k) = R^3 ij + (i+j) R^2 (R-1)/2 + (R-1)R(2R-1)/6  (SURVEY.md section 4);
    exact in fp64 while below 2^53."""
    i = np.asarray(i, dtype=np.float64)
    j = np.asarray(j, dtype=np.float64)
    return R ** 3 * i * j + (i + j) * (R * R * (R - 1) / 2.0) + (R - 1) * R * (2 * R - 1) / 6.0


# ---------------------------------------------------------------- global second oracle ----
def global_reference(rows, cols, vals, A, B):
    """scipy/numpy on the GLOBAL matrices: returns (S as scipy CSR with canonical order,
    sddmm values in that order, S@B, S.T@A, fused = (S o AB^T) @ B)."""
    import scipy.sparse as sp

    M, N = A.shape[0], B.shape[0]
    S = sp.csr_matrix((vals, (rows.astype(np.int64), cols.astype(np.int64))), shape=(M, N))
    S.sort_indices()
    ri = np.repeat(np.arange(M), np.diff(S.indptr))
    dots = np.einsum("ij,ij->i", A[ri], B[S.indices])
    sddmm = S.data * dots
    spmmA = S @ B
    spmmB = S.T @ A
    Sd = sp.csr_matrix((sddmm, S.indices, S.indptr), shape=(M, N))
    fused = Sd @ B
    return S, sddmm, spmmA, spmmB, fused


def gat_forward_global(rows, cols, N, layers, weights, alpha, X0):
    """GAT forward pass of the reference (gat.hpp:84-113) on GLOBAL matrices, for algorithms that do not split the
    dense operands along R.  Per head: H = X W; e = <H_u, H_v> on every stored (u, v) (the pattern values are 1.0,
    gat.hpp:87); e = max(e, 0) + alpha min(e, 0); Z = E H; out[:, head window] = max(Z, 0).  Pinned against the
    reference's own code in tests/test_oracle_vs_ref.py."""
    import scipy.sparse as sp

    S = sp.csr_matrix((np.ones(len(rows)), (rows.astype(np.int64), cols.astype(np.int64))), shape=(N, N))
    S.sort_indices()
    ri = np.repeat(np.arange(N), np.diff(S.indptr))
    X = np.asarray(X0, np.float64)
    for (fin, fph, heads), ws in zip(layers, weights):
        out = np.zeros((N, fph * heads))
        for h in range(heads):
            H = X @ np.asarray(ws[h], np.float64)
            e = np.einsum("ij,ij->i", H[ri], H[S.indices])
            e = np.maximum(e, 0.0) + np.minimum(e, 0.0) * alpha
            Z = sp.csr_matrix((e, S.indices, S.indptr), shape=(N, N)) @ H
            out[:, h * fph:(h + 1) * fph] = np.maximum(Z, 0.0)
        X = out
    return X
