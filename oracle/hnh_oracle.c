/*
 * oracle/hnh_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the local kernels of PASSIONLab/distributed_sddmm ("HnH").
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's library; the product (libhnh_b200.so) never links or
 * calls it and has no CPU fallback.
 *
 * PARITY STATUS of this restatement: the reference tree holds NO golden vectors,
 * known-answer tests or fixtures for this path, so by
 * itself this file is "parity unpinned".  It is pinned two ways in this repo:
 *   (1) oracle/_ref/ : the reference's own sparse_kernels.cpp / SpmatLocal.hpp /
 *       algorithm headers compiled from /root/reference against shim headers
 *       (oracle/shims/) -- see oracle/Makefile and tests/test_oracle_vs_ref.py;
 *   (2) scipy.sparse + closed-form dummyInitialize answers (tests/test_oracle.py).
 *
 * Every function cites the reference file:line it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* SDDMM over a block held in COO-expanded CSR order.
 * Follows sparse_kernels.cpp:44-55 literally: one nonzero per iteration, sequential
 * k-order fp64 dot product, "values[i] += dot".  (X,Y) are already role-resolved by the
 * caller: (A,B) for a non-transposed block, (B,A) for a transposed one
 * (sparse_kernels.cpp:29-38). */
void oracle_sddmm_coo(const int64_t *row_idx, const int64_t *col_idx, double *values,
                      int64_t nnz, const double *X, const double *Y, int64_t r)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nnz; i++) {
        const double *xr = X + r * row_idx[i];
        const double *yr = Y + r * col_idx[i];
        double acc = 0.0;
        for (int64_t k = 0; k < r; k++) acc += xr[k] * yr[k];
        values[i] += acc;
    }
}

/* SpMM  Y = 1*CSR*X + 1*Y, general, zero-based, row-major, ld = r.
 * Documented semantics of the mkl_sparse_d_mm call at sparse_kernels.cpp:95-120
 * (alpha = beta = 1, SPARSE_LAYOUT_ROW_MAJOR, ldb = ldc = R).  A row's nonzeros are
 * summed in stored CSR order. */
void oracle_spmm_csr(const int64_t *rowStart, const int64_t *col_idx, const double *values,
                     int64_t rows, const double *X, double *Y, int64_t r)
{
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < rows; i++) {
        double *yr = Y + r * i;
        for (int64_t j = rowStart[i]; j < rowStart[i + 1]; j++) {
            const double v = values[j];
            const double *xr = X + r * col_idx[j];
            for (int64_t k = 0; k < r; k++) yr[k] += v * xr[k];
        }
    }
}

/* COO -> CSR of a block, optionally transposing (CSRLocal ctor, SpmatLocal.hpp:78-188:
 * mkl_sparse_d_create_coo + mkl_sparse_convert_csr(op) + export).  Canonical order:
 * stored-row ascending, then input order (stable).  The reference feeds tuples sorted
 * (col, then row) (SpmatLocal.hpp:458), so stored-col is ascending inside a row for both
 * orientations.  Duplicates are kept (convert_csr does not merge).
 * out_rows = transpose ? cols : rows.  rowStart has out_rows+1 entries.  row_idx is the
 * expanded stored-row index per nonzero (SpmatLocal.hpp:139-147,160-163). */
int oracle_coo_to_csr(int64_t rows, int64_t cols, int64_t nnz,
                      const uint64_t *r, const uint64_t *c, const double *v, int transpose,
                      int64_t *rowStart, int64_t *col_idx, int64_t *row_idx, double *values)
{
    const int64_t out_rows = transpose ? cols : rows;
    const int64_t out_cols = transpose ? rows : cols;
    memset(rowStart, 0, sizeof(int64_t) * (size_t)(out_rows + 1));
    for (int64_t i = 0; i < nnz; i++) {
        int64_t sr = (int64_t)(transpose ? c[i] : r[i]);
        int64_t sc = (int64_t)(transpose ? r[i] : c[i]);
        if (sr < 0 || sr >= out_rows || sc < 0 || sc >= out_cols) return -1;
        rowStart[sr + 1]++;
    }
    for (int64_t i = 0; i < out_rows; i++) rowStart[i + 1] += rowStart[i];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(out_rows > 0 ? out_rows : 1));
    if (!cursor) return -2;
    memcpy(cursor, rowStart, sizeof(int64_t) * (size_t)out_rows);
    for (int64_t i = 0; i < nnz; i++) {
        int64_t sr = (int64_t)(transpose ? c[i] : r[i]);
        int64_t sc = (int64_t)(transpose ? r[i] : c[i]);
        int64_t p = cursor[sr]++;
        col_idx[p] = sc;
        row_idx[p] = sr;
        values[p] = v[i];
    }
    free(cursor);
    return 0;
}

/* Fused SDDMM->SpMM on one block = the two back-to-back triple_function calls of the
 * fusion-2 loop body (15D_dense_shift.hpp:203-217): K1 (values += X.Y) on the whole
 * block, then K2 (Out += CSR * Y) with the values just produced.  Two passes, exactly as
 * the reference executes them. */
void oracle_fused_block(const int64_t *rowStart, const int64_t *row_idx, const int64_t *col_idx,
                        double *values, int64_t rows, int64_t nnz,
                        const double *X, const double *Y, double *Out, int64_t r)
{
    oracle_sddmm_coo(row_idx, col_idx, values, nnz, X, Y, r);
    oracle_spmm_csr(rowStart, col_idx, values, rows, Y, Out, r);
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
