// oracle/shims/mkl_spblas.h -- TEST INFRASTRUCTURE.  Restatement of the six Intel MKL
// inspector-executor routines the reference calls (SpmatLocal.hpp:117-192,
// sparse_kernels.cpp:95-120), so that the reference's own sources compile and run UNMODIFIED
// in oracle/_ref.  MKL is a closed third-party binary that is not in this image and not pinned
// by the reference (README.md:72 says ">= 2018"); what is restated is its DOCUMENTED behaviour:
//   * mkl_sparse_d_create_coo / create_csr wrap the caller's arrays (no copy);
//   * mkl_sparse_convert_csr(op): COO -> CSR of A (op = NON_TRANSPOSE) or of A^T (TRANSPOSE),
//     entries of a row kept in input order (stable), duplicates kept;
//   * mkl_sparse_d_export_csr hands out the internal 4-array CSR;
//   * mkl_sparse_d_mm: Y = alpha * op(A) * X + beta * Y, row-major dense operands.
// Anything the reference does not use returns SPARSE_STATUS_NOT_SUPPORTED.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef MKL_INT
#define MKL_INT long long
#endif

typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20 } sparse_matrix_type_t;
typedef enum { SPARSE_LAYOUT_ROW_MAJOR = 101, SPARSE_LAYOUT_COLUMN_MAJOR = 102 } sparse_layout_t;
struct matrix_descr {
    sparse_matrix_type_t type;
    int mode, diag;
};

struct hnh_shim_sparse_matrix {
    bool is_coo;
    MKL_INT rows, cols, nnz;
    // views (create_*) or owned storage (convert_csr)
    MKL_INT *coo_r, *coo_c;
    MKL_INT *rows_start, *rows_end, *col_idx;
    double *values;
    std::vector<MKL_INT> own_start, own_col;
    std::vector<double> own_val;
};
typedef hnh_shim_sparse_matrix *sparse_matrix_t;

inline sparse_status_t mkl_sparse_d_create_coo(sparse_matrix_t *A, sparse_index_base_t, MKL_INT rows, MKL_INT cols,
                                               MKL_INT nnz, MKL_INT *row_indx, MKL_INT *col_indx, double *values) {
    hnh_shim_sparse_matrix *m = new hnh_shim_sparse_matrix();
    m->is_coo = true; m->rows = rows; m->cols = cols; m->nnz = nnz;
    m->coo_r = row_indx; m->coo_c = col_indx; m->values = values;
    m->rows_start = m->rows_end = m->col_idx = nullptr;
    *A = m;
    return SPARSE_STATUS_SUCCESS;
}

inline sparse_status_t mkl_sparse_convert_csr(const sparse_matrix_t src, sparse_operation_t op, sparse_matrix_t *dst) {
    if (!src->is_coo) return SPARSE_STATUS_NOT_SUPPORTED;
    const bool t = op == SPARSE_OPERATION_TRANSPOSE;
    hnh_shim_sparse_matrix *m = new hnh_shim_sparse_matrix();
    m->is_coo = false;
    m->rows = t ? src->cols : src->rows;
    m->cols = t ? src->rows : src->cols;
    m->nnz = src->nnz;
    m->own_start.assign((size_t)m->rows + 1, 0);
    m->own_col.resize((size_t)m->nnz);
    m->own_val.resize((size_t)m->nnz);
    const MKL_INT *sr = t ? src->coo_c : src->coo_r, *sc = t ? src->coo_r : src->coo_c;
    for (MKL_INT i = 0; i < m->nnz; i++) m->own_start[(size_t)sr[i] + 1]++;
    for (MKL_INT i = 0; i < m->rows; i++) m->own_start[(size_t)i + 1] += m->own_start[(size_t)i];
    std::vector<MKL_INT> cursor(m->own_start.begin(), m->own_start.end() - 1);
    for (MKL_INT i = 0; i < m->nnz; i++) {
        const MKL_INT p = cursor[(size_t)sr[i]]++;
        m->own_col[(size_t)p] = sc[i];
        m->own_val[(size_t)p] = src->values[i];
    }
    m->rows_start = m->own_start.data();
    m->rows_end = m->own_start.data() + 1;
    m->col_idx = m->own_col.data();
    m->values = m->own_val.data();
    m->coo_r = m->coo_c = nullptr;
    *dst = m;
    return SPARSE_STATUS_SUCCESS;
}

inline sparse_status_t mkl_sparse_d_export_csr(const sparse_matrix_t A, sparse_index_base_t *indexing, MKL_INT *rows,
                                               MKL_INT *cols, MKL_INT **rows_start, MKL_INT **rows_end, MKL_INT **col_indx,
                                               double **values) {
    if (A->is_coo) return SPARSE_STATUS_NOT_SUPPORTED;
    *indexing = SPARSE_INDEX_BASE_ZERO;
    *rows = A->rows; *cols = A->cols;
    *rows_start = A->rows_start; *rows_end = A->rows_end; *col_indx = A->col_idx; *values = A->values;
    return SPARSE_STATUS_SUCCESS;
}

inline sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, sparse_index_base_t, MKL_INT rows, MKL_INT cols,
                                               MKL_INT *rows_start, MKL_INT *rows_end, MKL_INT *col_indx, double *values) {
    hnh_shim_sparse_matrix *m = new hnh_shim_sparse_matrix();
    m->is_coo = false; m->rows = rows; m->cols = cols; m->nnz = -1;
    m->rows_start = rows_start; m->rows_end = rows_end; m->col_idx = col_indx; m->values = values;
    m->coo_r = m->coo_c = nullptr;
    *A = m;
    return SPARSE_STATUS_SUCCESS;
}

inline sparse_status_t mkl_sparse_destroy(sparse_matrix_t A) {
    delete A;
    return SPARSE_STATUS_SUCCESS;
}

// Y = alpha * A * X + beta * Y ; row-major X (cols(A) x columns, ld ldx), Y (rows(A) x columns, ld ldy).
// Reads A's arrays at call time (the reference relies on that: it rewrites values / rowStart in
// place between calls, SpmatLocal.hpp:571-605,200-259).
inline sparse_status_t mkl_sparse_d_mm(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr,
                                       sparse_layout_t layout, const double *x, MKL_INT columns, MKL_INT ldx, double beta,
                                       double *y, MKL_INT ldy) {
    if (A->is_coo || op != SPARSE_OPERATION_NON_TRANSPOSE || layout != SPARSE_LAYOUT_ROW_MAJOR)
        return SPARSE_STATUS_NOT_SUPPORTED;
    #pragma omp parallel for schedule(dynamic, 256)
    for (MKL_INT i = 0; i < A->rows; i++) {
        double *yr = y + i * ldy;
        if (beta != 1.0)
            for (MKL_INT k = 0; k < columns; k++) yr[k] *= beta;
        for (MKL_INT j = A->rows_start[i]; j < A->rows_end[i]; j++) {
            const double v = alpha * A->values[j];
            const double *xr = x + A->col_idx[j] * ldx;
            for (MKL_INT k = 0; k < columns; k++) yr[k] += v * xr[k];
        }
    }
    return SPARSE_STATUS_SUCCESS;
}
