"""ctypes front end of oracle/_ref/libhnh_ref.so -- TEST INFRASTRUCTURE.

libhnh_ref.so is the REFERENCE's own code (PASSIONLab/distributed_sddmm, compiled unmodified
from /root/reference by oracle/ref.mk) running with MPI ranks as threads on shimmed MPI / MKL /
Eigen / CombBLAS.  It is the strongest checker this repo has: every distributed layout, block
split, value order and ring data-flow in it is the reference's actual code.  Built only where
/root/reference exists (this container); the built .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libhnh_ref.so")
# the same reference code with this repo's CUDA kernels plugged in through KernelImplementation (ref.mk `cuda`)
SO_CUDA = os.path.join(_HERE, "_ref", "libhnh_ref_cuda.so")
_LIB = None
_LIB_CUDA = None


def available() -> bool:
    return os.path.exists(SO)


def build() -> bool:
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", _HERE, "-f", "ref.mk"], stdout=subprocess.DEVNULL)
        build_cuda_plugin()
    return available()


def build_cuda_plugin() -> bool:
    """Optional second build: needs the product library (distributed_sddmm_b200/libhnh_b200.so) to link against."""
    if os.path.isdir("/root/reference"):
        subprocess.call(["make", "-C", _HERE, "-f", "ref.mk", "cuda"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(SO_CUDA)


def lib(plugin: bool = False):
    global _LIB, _LIB_CUDA
    if plugin:
        if _LIB_CUDA is None:
            _LIB_CUDA = _bind(C.CDLL(SO_CUDA))
        return _LIB_CUDA
    if _LIB is None:
        _LIB = _bind(C.CDLL(SO))
    return _LIB


def _bind(L):
    P = C.c_void_p
    L.ref_run.restype = P
    L.ref_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P,
                          C.c_char_p, C.c_int]
    L.ref_error.restype = C.c_char_p
    L.ref_error.argtypes = [P]
    L.ref_free.argtypes = [P]
    L.ref_rank_info.argtypes = [P, C.c_int, P]
    for name in ("ref_submatrices",):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [P, C.c_int, C.c_int]
    L.ref_num_values.restype = C.c_int64
    L.ref_num_values.argtypes = [P, C.c_int, C.c_int]
    for name in ("ref_value_rows", "ref_value_cols"):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [P, C.c_int, C.c_int]
    L.ref_block_meta.argtypes = [P, C.c_int, C.c_int, C.c_int, P]
    for name in ("ref_block_rowStart", "ref_block_col_idx", "ref_block_row_idx", "ref_block_values"):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [P, C.c_int, C.c_int, C.c_int]
    for name in ("ref_op_A", "ref_op_B", "ref_op_values"):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [P, C.c_int, C.c_int]
    L.ref_op_num_values.restype = C.c_int64
    L.ref_op_num_values.argtypes = [P, C.c_int, C.c_int]
    L.ref_op_elapsed.restype = C.c_double
    L.ref_op_elapsed.argtypes = [P, C.c_int, C.c_int]
    L.ref_benchmark.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int,
                                C.c_char_p, C.c_char_p, C.c_int]
    L.ref_benchmark.restype = None
    L.ref_time_fused.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                 C.c_int, P]
    L.ref_time_fused.restype = C.c_int64
    if hasattr(L, "ref_time_fused_check"):
        L.ref_time_fused_check.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                           C.c_int, P, P]
        L.ref_time_fused_check.restype = C.c_int64
    return L


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double if dtype == np.float64 else
                                                         C.c_int64 if dtype == np.int64 else C.c_int32)),
                                 shape=(n,)).copy()


def run(alg: str, p: int, c: int, R: int, M: int, N: int, rows, cols, vals, A, B, script, threads_per_rank: int = 1,
        plugin: bool = False):
    """Run `script` (list of ops) of the reference's `alg` on p thread-ranks.  rows/cols/vals must be
    sorted by (row, col).  Returns a list (one per rank) of dicts with the layout, the local CSR
    blocks, the global coordinates of every local value slot and the per-op local outputs.
    plugin=True: the build whose local kernels are this repo's CUDA kernels behind the reference's
    KernelImplementation interface (needs a GPU)."""
    L = lib(plugin)
    rows = np.ascontiguousarray(rows, np.uint64)
    cols = np.ascontiguousarray(cols, np.uint64)
    vals = np.ascontiguousarray(vals, np.float64)
    A = np.ascontiguousarray(A, np.float64)
    B = np.ascontiguousarray(B, np.float64)
    h = L.ref_run(alg.encode(), p, c, R, M, N, len(rows), rows.ctypes.data, cols.ctypes.data, vals.ctypes.data,
                  A.ctypes.data, B.ctypes.data, ",".join(script).encode(), threads_per_rank)
    try:
        err = L.ref_error(h).decode()
        if err:
            raise RuntimeError(err)
        out = []
        for rank in range(p):
            info = (C.c_int * 11)()
            L.ref_rank_info(h, rank, info)
            i, j, k, lAr, lAc, lBr, lBc, na, nb, nsb, nstb = list(info)
            d = dict(i=i, j=j, k=k, localArows=lAr, localAcols=lAc, localBrows=lBr, localBcols=lBc)
            d["aSubmatrices"] = _arr(L.ref_submatrices(h, rank, 0), 4 * na, np.int32).reshape(na, 4)
            d["bSubmatrices"] = _arr(L.ref_submatrices(h, rank, 1), 4 * nb, np.int32).reshape(nb, 4)
            for w, key, nblk in ((0, "S", nsb), (1, "ST", nstb)):
                n = L.ref_num_values(h, rank, w)
                d[key + "_rows"] = _arr(L.ref_value_rows(h, rank, w), n, np.int64)
                d[key + "_cols"] = _arr(L.ref_value_cols(h, rank, w), n, np.int64)
                blocks = []
                for b in range(nblk):
                    meta = (C.c_int64 * 5)()
                    L.ref_block_meta(h, rank, w, b, meta)
                    if meta[0]:
                        blocks.append(None)
                        continue
                    nr, nc, nnz, tr = meta[1], meta[2], meta[3], meta[4]
                    blocks.append(dict(rows=nr, cols=nc, transpose=bool(tr),
                                       rowStart=_arr(L.ref_block_rowStart(h, rank, w, b), nr + 1, np.int64),
                                       col_idx=_arr(L.ref_block_col_idx(h, rank, w, b), nnz, np.int64),
                                       row_idx=_arr(L.ref_block_row_idx(h, rank, w, b), nnz, np.int64),
                                       values=_arr(L.ref_block_values(h, rank, w, b), nnz, np.float64)))
                d[key + "_blocks"] = blocks
            ops = []
            for t, name in enumerate(script):
                nv = L.ref_op_num_values(h, rank, t)
                ops.append(dict(op=name, A=_arr(L.ref_op_A(h, rank, t), lAr * lAc, np.float64).reshape(lAr, lAc),
                                B=_arr(L.ref_op_B(h, rank, t), lBr * lBc, np.float64).reshape(lBr, lBc),
                                values=_arr(L.ref_op_values(h, rank, t), nv, np.float64),
                                elapsed=L.ref_op_elapsed(h, rank, t)))
            d["ops"] = ops
            out.append(d)
        return out
    finally:
        L.ref_free(h)


def assemble_dense(ranks, which: str, op_index: int, rows: int, R: int):
    """Global matrix from the per-rank local outputs via the submatrix descriptors (replicas agree)."""
    G = np.zeros((rows, R))
    for d in ranks:
        loc = d["ops"][op_index][which]
        subs = d["aSubmatrices" if which == "A" else "bSubmatrices"]
        at = 0
        flat = loc.reshape(-1)
        for top, left, nr, nc in subs:
            blk = flat[at:at + nr * nc].reshape(nr, nc)
            at += nr * nc
            r_hi = min(top + nr, rows)
            if r_hi > top:
                G[top:r_hi, left:left + nc] = blk[:r_hi - top]
    return G


def benchmark(alg, p, c, R, logM, nnz_per_row, seed, fused=True, app="vanilla", output_file="/tmp/ref_bench.json",
              threads_per_rank=1):
    lib().ref_benchmark(alg.encode(), p, c, R, logM, nnz_per_row, seed, int(fused), app.encode(),
                        output_file.encode(), threads_per_rank)


def time_fused(alg, p, c, R, logM, nnz_per_row, seed, warmup, steps, threads_per_rank):
    """Seconds of each of (warmup + steps) fusedSpMM(A, B, S, result, Amat) calls of the REFERENCE
    (its own loadTuples -> algorithm -> kernels), benchmark inputs.  Returns (dist_nnz, seconds[])."""
    secs = np.zeros(warmup + steps)
    nnz = lib().ref_time_fused(alg.encode(), p, c, R, logM, nnz_per_row, seed, warmup, steps, threads_per_rank,
                               secs.ctypes.data)
    return int(nnz), secs


def pattern(rows, cols, salt, row0=0, col0=0):
    """The position-dependent test operand of the parity legs (ref_driver.cpp::pattern_value): global rows
    [row0, row0 + rows) x columns [col0, col0 + cols).  Exactly representable in fp64."""
    return pattern_rows(np.arange(row0, row0 + rows, dtype=np.uint64), cols, salt, col0)


def pattern_rows(row_index, cols, salt, col0=0):
    """pattern() for an arbitrary list of global row indices."""
    i = np.asarray(row_index, dtype=np.uint64)[:, None]
    k = np.arange(col0, col0 + cols, dtype=np.uint64)[None, :]
    h = (i * np.uint64(2654435761) + k * np.uint64(2246822519) + np.uint64(salt * 97)) & np.uint64(0xFFFFFFFF)
    return h.astype(np.float64) / 4294967296.0 - 0.5


def time_fused_check(alg, c, R, logM, nnz_per_row, seed, warmup, steps, threads):
    """time_fused on ONE thread-rank plus the parity leg: one more fusedSpMM of the reference on A = pattern(1),
    B = pattern(2).  Returns (dist_nnz, seconds[], result N x R)."""
    secs = np.zeros(warmup + steps)
    out = np.empty(((1 << logM), R))
    nnz = lib().ref_time_fused_check(alg.encode(), c, R, logM, nnz_per_row, seed, warmup, steps, threads,
                                     secs.ctypes.data, out.ctypes.data)
    return int(nnz), secs, out


def gat(alg, p, c, N, rows, cols, vals, layers, weights, alpha, X0, threads_per_rank: int = 1):
    """The REFERENCE's GAT forward pass (gat.hpp) on p thread-ranks.  layers: [(in, per_head, heads)];
    weights[i][h]: in x per_head; X0: N x layers[0][0].  Returns (global output N x out_width, per-rank list of
    (local buffer, aSubmatrices))."""
    L = lib()
    P = C.c_void_p
    L.ref_gat.restype = P
    L.ref_gat.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, C.c_int, P, P,
                          C.c_double, P, C.c_int]
    L.ref_gat_error.restype = C.c_char_p
    L.ref_gat_error.argtypes = [P]
    L.ref_gat_shape_len.argtypes = [P, C.c_int]
    L.ref_gat_shape.restype = P
    L.ref_gat_shape.argtypes = [P, C.c_int]
    L.ref_gat_out.restype = P
    L.ref_gat_out.argtypes = [P, C.c_int]
    L.ref_gat_free.argtypes = [P]
    rows = np.ascontiguousarray(rows, np.uint64)
    cols = np.ascontiguousarray(cols, np.uint64)
    vals = np.ascontiguousarray(vals, np.float64)
    l3 = np.ascontiguousarray(np.array(layers, dtype=np.int32).reshape(-1))
    w = np.ascontiguousarray(np.concatenate([np.asarray(weights[i][h], np.float64).reshape(-1)
                                             for i in range(len(layers)) for h in range(layers[i][2])]))
    X0 = np.ascontiguousarray(X0, np.float64)
    h = L.ref_gat(alg.encode(), p, c, N, N, len(rows), rows.ctypes.data, cols.ctypes.data, vals.ctypes.data, len(layers),
                  l3.ctypes.data, w.ctypes.data, float(alpha), X0.ctypes.data, threads_per_rank)
    try:
        err = L.ref_gat_error(h).decode()
        if err:
            raise RuntimeError(err)
        outw = layers[-1][1] * layers[-1][2]
        G = np.zeros((N, outw))
        per_rank = []
        for rank in range(p):
            n = L.ref_gat_shape_len(h, rank)
            sh = _arr(L.ref_gat_shape(h, rank), n, np.int32)
            nr, nc = int(sh[0]), int(sh[1])
            subs = sh[2:].reshape(-1, 4)
            loc = _arr(L.ref_gat_out(h, rank), nr * nc, np.float64).reshape(nr, nc)
            per_rank.append((loc, subs))
            at, flat = 0, loc.reshape(-1)
            for top, left, snr, snc in subs:
                blk = flat[at:at + snr * snc].reshape(snr, snc)
                at += snr * snc
                hi = min(top + snr, N)
                if hi > top:
                    G[top:hi, left:left + snc] = blk[:hi - top]
        return G, per_rank
    finally:
        L.ref_gat_free(h)


def als(alg, p, c, R, N, rows, cols, Agt, Bgt, A0, B0, steps: int = 1, cg_iters: int = 10, threads_per_rank: int = 1):
    """The REFERENCE's Distributed_ALS on caller inputs (global N x R matrices; see ref_driver.cpp::ref_als).
    Returns dict(residual=(before, after), A=global A, B=global B, ranks=[(localA, localB)])."""
    L = lib()
    P = C.c_void_p
    L.ref_als.restype = P
    L.ref_als.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P, P, P, C.c_int,
                          C.c_int, C.c_int]
    L.ref_als_error.restype = C.c_char_p
    L.ref_als_error.argtypes = [P]
    L.ref_als_residuals.argtypes = [P, P]
    L.ref_als_shape_len.argtypes = [P, C.c_int]
    for name in ("ref_als_shape", "ref_als_A", "ref_als_B"):
        getattr(L, name).restype = P
        getattr(L, name).argtypes = [P, C.c_int]
    L.ref_als_free.argtypes = [P]
    rows = np.ascontiguousarray(rows, np.uint64)
    cols = np.ascontiguousarray(cols, np.uint64)
    vals = np.ones(len(rows))
    mats = [np.ascontiguousarray(m, np.float64) for m in (Agt, Bgt, A0, B0)]
    h = L.ref_als(alg.encode(), p, c, R, N, N, len(rows), rows.ctypes.data, cols.ctypes.data, vals.ctypes.data,
                  *[m.ctypes.data for m in mats], steps, cg_iters, threads_per_rank)
    try:
        err = L.ref_als_error(h).decode()
        if err:
            raise RuntimeError(err)
        res = (C.c_double * 2)()
        L.ref_als_residuals(h, res)
        GA, GB = np.zeros((N, R)), np.zeros((N, R))
        ranks = []
        for rank in range(p):
            sh = _arr(L.ref_als_shape(h, rank), L.ref_als_shape_len(h, rank), np.int32)
            lAr, lAc, lBr, lBc, na, nb = (int(x) for x in sh[:6])
            subsA = sh[6:6 + 4 * na].reshape(na, 4)
            subsB = sh[6 + 4 * na:6 + 4 * (na + nb)].reshape(nb, 4)
            locA = _arr(L.ref_als_A(h, rank), lAr * lAc, np.float64).reshape(lAr, lAc)
            locB = _arr(L.ref_als_B(h, rank), lBr * lBc, np.float64).reshape(lBr, lBc)
            ranks.append((locA, locB))
            for G, loc, subs in ((GA, locA, subsA), (GB, locB, subsB)):
                at, flat = 0, loc.reshape(-1)
                for top, left, nr, nc in subs:
                    blk = flat[at:at + nr * nc].reshape(nr, nc)
                    at += nr * nc
                    hi = min(top + nr, N)
                    if hi > top:
                        G[top:hi, left:left + nc] = blk[:hi - top]
        return dict(residual=(res[0], res[1]), A=GA, B=GB, ranks=ranks)
    finally:
        L.ref_als_free(h)
