// hnh/sparse_kernels.h -- the local-kernel plugin interface of the reference
// (sparse_kernels.h:13-99: KernelMode, KernelImplementation, StandardKernel), with
// StandardKernel implemented by the sm_100a CUDA kernels behind the C ABI of hnh_b200.h
// instead of an OpenMP loop + Intel MKL.
//
// Contract kept from the reference: sddmm_local ACCUMULATES into the block's CSR values,
// spmm_local ACCUMULATES into the output matrix (callers zero first), a null block or an empty
// block is a silent no-op, both return 0 ("nnz processed" is never counted,
// sparse_kernels.cpp:23,56,66,126).  Difference: a transposition / mode mismatch throws
// hnh::Error(HNH_E_MODE) instead of printing and calling exit(1) (sparse_kernels.cpp:75-83).
// Work is ENQUEUED on hnh::Runtime::compute_stream(); every consumer in this library is
// stream-ordered behind it.
#pragma once
#include <cstddef>

#include "hnh/SpmatLocal.hpp"
#include "hnh/common.h"

typedef enum { k_sddmmA, k_spmmA, k_spmmB, k_sddmmB } KernelMode;

class KernelImplementation {
public:
    virtual ~KernelImplementation() {}

    // values[i] += X[row_idx[i]] . Y[col_idx[i]] on S.csr_blocks[block]; (X, Y) = (A, B) for a
    // non-transposed block, (B, A) for a transposed one.  `offset` is ignored (as in the
    // reference).
    virtual size_t sddmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, int block, int offset) = 0;

    // mode == Amat: A += S B (needs a non-transposed block); mode == Bmat: B += S_stored A
    // (needs a transposed block).
    virtual size_t spmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, MatMode mode, int block) = 0;

    // SDDMM immediately followed by SpMM on the same block with the values just produced (the
    // fusion-2 loop body, 15D_dense_shift.hpp:203-217): values += X.B rows, Out += S B.
    // The default is the reference's two calls; StandardKernel overrides it with ONE kernel that
    // gathers each B row once.  `first_visit`: the block's values are known to be zero and, if
    // `out_is_zero`, so is Out (lets the CUDA kernel skip reading them).
    virtual size_t fused_local(SpmatLocal &S, DenseMatrix &X, DenseMatrix &B, DenseMatrix &Out, int block,
                               bool first_visit, bool out_is_zero) {
        (void)first_visit;
        (void)out_is_zero;
        size_t n = sddmm_local(S, X, B, block, 0);
        n += spmm_local(S, Out, B, Amat, block);
        return n;
    }

    size_t triple_function(KernelMode mode, SpmatLocal &S, DenseMatrix &localA, DenseMatrix &localB, int block,
                           int offset) {
        switch (mode) {
            case k_sddmmA:
            case k_sddmmB: return sddmm_local(S, localA, localB, block, offset);
            case k_spmmA: return spmm_local(S, localA, localB, Amat, block);
            case k_spmmB: return spmm_local(S, localA, localB, Bmat, block);
        }
        return 0;
    }
};

// "Exactly the algebra on the box" -- on the B200.
class StandardKernel : public KernelImplementation {
public:
    // hints set by the algorithm classes around a call (never change results):
    bool values_are_zero = false;  // next sddmm_local may overwrite instead of accumulate
    bool output_is_zero = false;   // next spmm_local may overwrite instead of accumulate
    int flags = 0;                 // extra HNH_FLAG_* bits (testing)
    // SDDMM epilogue (hnh_sddmm_scaled_f64): while sddmm_scale is set, sddmm_local multiplies the finished dot of
    // nonzero i of block b by (*sddmm_scale)[S.blockStarts[b] + i] -- the reference's separate
    // `SValues.cwiseProduct(getCSRValues())` pass (15D_dense_shift.hpp:364-368) -- and stores the product in
    // *sddmm_scaled_out (same indexing) and, with sddmm_scale_values, in the block's CSR values as well.  Only legal
    // when the call completes the dot (every block is met once per pass, as in the 1.5D dense-shift algorithm).
    const VectorXd *sddmm_scale = nullptr;
    VectorXd *sddmm_scaled_out = nullptr;
    bool sddmm_scale_values = false;

    size_t sddmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, int block, int offset) override;
    size_t spmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, MatMode mode, int block) override;
    size_t fused_local(SpmatLocal &S, DenseMatrix &X, DenseMatrix &B, DenseMatrix &Out, int block, bool first_visit,
                       bool out_is_zero) override;
    // fused_local restricted to the CSR rows [row0, row0 + nrows) of the block: the unit of the host-operand
    // pipeline (fusedSpMM_host).  The block's values of those rows are overwritten (first visit); the rows of
    // Out are overwritten when out_is_zero (Out may then be X), accumulated into otherwise.  Only for the widths
    // of the dispatch table (4..256, powers of two); throws otherwise.
    void fused_local_rows(SpmatLocal &S, DenseMatrix &X, DenseMatrix &B, DenseMatrix &Out, int block, int64_t row0,
                          int64_t nrows, bool out_is_zero = true);
};
