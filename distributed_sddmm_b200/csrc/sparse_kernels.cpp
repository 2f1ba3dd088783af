// sparse_kernels.cpp -- StandardKernel on the sm_100a kernels (hnh_b200.h C ABI).
#include "hnh/sparse_kernels.h"

#include "hnh_b200.h"

using hnh::abi_check;
using hnh::Runtime;

size_t StandardKernel::sddmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, int block, int /*offset*/) {
    if (A.cols() != B.cols()) throw hnh::Error(HNH_E_INVALID, "sddmm_local: A.cols() != B.cols()");
    CSRLocal *blk = S.csr_blocks[block];
    if (blk == nullptr || blk->num_coords == 0) return 0;
    // role swap on transposed storage (reference sparse_kernels.cpp:29-38)
    const double *X = blk->transpose ? B.data() : A.data();
    const double *Y = blk->transpose ? A.data() : B.data();
    CSRHandle *h = blk->getActive();
    int f = flags | (values_are_zero ? HNH_FLAG_BETA0 : 0);
    const double *scale = nullptr;
    double *scaled_out = nullptr;
    if (sddmm_scale != nullptr) {
        const int64_t off = (int64_t)S.blockStarts[(size_t)block];
        if (off + blk->num_coords > sddmm_scale->size() || (sddmm_scaled_out && off + blk->num_coords > sddmm_scaled_out->size()))
            throw hnh::Error(HNH_E_INVALID, "sddmm_local: the S value vectors are shorter than the local sparse matrix");
        scale = sddmm_scale->data() + off;
        if (sddmm_scaled_out) scaled_out = sddmm_scaled_out->data() + off;
        if (sddmm_scale_values) f |= HNH_FLAG_SCALE_VALUES;
    }
    abi_check(hnh_sddmm_scaled_f64(h->rowStart.data(), h->col_idx.data(), h->values.data(), blk->rows, blk->num_coords, X, Y,
                                   (int)A.cols(), f, scale, scaled_out, Runtime::get().compute_stream()),
              "hnh_sddmm_f64");
    return 0;
}

size_t StandardKernel::spmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, MatMode mode, int block) {
    CSRLocal *blk = S.csr_blocks[block];
    if (output_is_zero && (blk == nullptr || blk->num_coords == 0)) {
        // "overwrite" hint on an empty block: the output of this step is all zeros
        ((mode == Amat) ? A : B).setZero();
        return 0;
    }
    if (blk == nullptr) return 0;
    if (mode == Amat && blk->transpose)
        throw hnh::Error(HNH_E_MODE, "Error, local matrix is transposed, can't perform SpmmA");
    if (mode == Bmat && !blk->transpose)
        throw hnh::Error(HNH_E_MODE, "Error, local matrix is not transposed, can't perform SpmmB");
    if (blk->num_coords == 0) return 0;
    const double *In = (mode == Amat) ? B.data() : A.data();
    double *Out = (mode == Amat) ? A.data() : B.data();
    CSRHandle *h = blk->getActive();
    int f = flags | (output_is_zero ? HNH_FLAG_BETA0 : 0);
    abi_check(hnh_spmm_f64(h->rowStart.data(), h->col_idx.data(), h->values.data(), blk->rows, blk->num_coords, In, Out,
                           (int)A.cols(), f, Runtime::get().compute_stream()),
              "hnh_spmm_f64");
    return 0;
}

void StandardKernel::fused_local_rows(SpmatLocal &S, DenseMatrix &X, DenseMatrix &B, DenseMatrix &Out, int block,
                                      int64_t row0, int64_t nrows, bool out_is_zero) {
    CSRLocal *blk = S.csr_blocks[block];
    if (nrows <= 0) return;
    const int64_t r = X.cols();
    if (blk == nullptr || blk->num_coords == 0) {
        if (out_is_zero)
            hnh::cuda_check(cudaMemsetAsync(Out.data() + row0 * r, 0, sizeof(double) * (size_t)(nrows * r),
                                            Runtime::get().compute_stream()), "cudaMemsetAsync");
        return;
    }
    if (blk->transpose) throw hnh::Error(HNH_E_MODE, "fused_local_rows needs a non-transposed block");
    if (r < 4 || r > 256 || (r & (r - 1)) != 0)  // the generic kernel's overwrite mode clears the whole block's values
        throw hnh::Error(HNH_E_INVALID, "fused_local_rows: width outside the dispatch table");
    if (row0 < 0 || row0 + nrows > blk->rows) throw hnh::Error(HNH_E_INVALID, "fused_local_rows: row range");
    CSRHandle *h = blk->getActive();
    // rowStart entries are absolute offsets into col_idx / values, so a row range is the same call on shifted
    // rowStart / X / Out pointers
    abi_check(hnh_fused_f64(h->rowStart.data() + row0, h->col_idx.data(), h->values.data(), nrows, blk->num_coords,
                            X.data() + row0 * r, B.data(), Out.data() + row0 * r, (int)r,
                            flags | HNH_FLAG_BETA0_VALUES | (out_is_zero ? HNH_FLAG_BETA0_OUT : 0),
                            Runtime::get().compute_stream()),
              "hnh_fused_f64");
}

size_t StandardKernel::fused_local(SpmatLocal &S, DenseMatrix &X, DenseMatrix &B, DenseMatrix &Out, int block,
                                   bool first_visit, bool out_is_zero) {
    CSRLocal *blk = S.csr_blocks[block];
    if (blk == nullptr || blk->num_coords == 0) {
        if (out_is_zero) Out.setZero();
        return 0;
    }
    if (blk->transpose) throw hnh::Error(HNH_E_MODE, "fused_local needs a non-transposed block");
    CSRHandle *h = blk->getActive();
    int f = flags | (first_visit ? HNH_FLAG_BETA0_VALUES : 0) | (out_is_zero ? HNH_FLAG_BETA0_OUT : 0);
    const double *scale = nullptr;
    double *scaled_out = nullptr;
    if (sddmm_scale != nullptr) {  // S values folded into the fused kernel (see sparse_kernels.h)
        const int64_t off = (int64_t)S.blockStarts[(size_t)block];
        if (off + blk->num_coords > sddmm_scale->size() || (sddmm_scaled_out && off + blk->num_coords > sddmm_scaled_out->size()))
            throw hnh::Error(HNH_E_INVALID, "fused_local: the S value vectors are shorter than the local sparse matrix");
        scale = sddmm_scale->data() + off;
        if (sddmm_scaled_out) scaled_out = sddmm_scaled_out->data() + off;
    }
    abi_check(hnh_fused_scaled_f64(h->rowStart.data(), h->col_idx.data(), h->values.data(), blk->rows, blk->num_coords,
                                   X.data(), B.data(), Out.data(), (int)X.cols(), f, scale, scaled_out,
                                   Runtime::get().compute_stream()),
              "hnh_fused_f64");
    return 0;
}
