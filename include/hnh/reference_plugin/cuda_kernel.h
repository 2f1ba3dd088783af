// cuda_kernel.h -- the REFERENCE-SIDE binding of include/hnh_b200.h (INTEGRATION.md route B).
//
// This header is meant to be dropped into the reference tree (PASSIONLab/distributed_sddmm): it includes the
// reference's own "sparse_kernels.h" and derives `CudaKernel` from its plugin interface
// `KernelImplementation` (sparse_kernels.h:15-79).  Everything else of the reference -- Eigen matrices in host
// memory, MKL_INT index vectors, MPI, the 1.5D / 2.5D algorithm classes -- stays untouched; only
//     StandardKernel local_ops;   ->   CudaKernel local_ops;          (benchmark_dist.cpp:42)
// changes, and the program links libhnh_b200.so.  No CUDA toolchain is needed on the reference side: the two
// calls below take HOST pointers.  Every call mirrors its operands on the GPU and copies the result back, so this
// route is correct for every caller but PCIe-bound; route A (include/hnh/*.hpp) keeps the operands in HBM.
//
// Semantics kept (sparse_kernels.cpp:13-127): role swap of (A, B) on transposed blocks; SDDMM accumulates into the
// block's values and is driven by (row_idx, col_idx) -- the arrays the reference keeps current in SDDMM passes
// (SpmatLocal.hpp:223-240); SpMM is Y += CSR * X driven by (rowStart, col_idx), Amat needs a non-transposed
// block and Bmat a transposed one; null blocks are skipped; both return 0.  Errors throw instead of exit(1).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

#include "hnh_b200.h"
#include "sparse_kernels.h"  // the reference's

class CudaKernel : public KernelImplementation {
    static_assert(sizeof(MKL_INT) == sizeof(int64_t), "the B200 kernels take 64-bit indices (MKL ILP64)");
    static const int64_t *idx(const std::vector<MKL_INT> &v) { return reinterpret_cast<const int64_t *>(v.data()); }
    static void ok(int rc, const char *what) {
        if (rc != HNH_OK) throw std::runtime_error(std::string(what) + ": " + hnh_last_error_string());
    }

public:
    size_t sddmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, int block, int /*offset*/) override {
        CSRLocal *blk = S.csr_blocks[block];
        if (blk == nullptr) return 0;
        if (A.cols() != B.cols()) throw std::runtime_error("CudaKernel::sddmm_local: A and B differ in width");
        CSRHandle *h = blk->getActive();
        DenseMatrix &X = blk->transpose ? B : A;  // indexed by the stored row
        DenseMatrix &Y = blk->transpose ? A : B;  // gathered by the stored column
        ok(hnh_sddmm_coo_host(idx(h->row_idx), idx(h->col_idx), h->values.data(), (int64_t)blk->num_coords, X.data(),
                              (int64_t)X.rows(), Y.data(), (int64_t)Y.rows(), (int)A.cols()),
           "hnh_sddmm_coo_host");
        return 0;
    }

    size_t spmm_local(SpmatLocal &S, DenseMatrix &A, DenseMatrix &B, MatMode mode, int block) override {
        CSRLocal *blk = S.csr_blocks[block];
        if (blk == nullptr || blk->num_coords == 0) return 0;
        if (mode == Amat && blk->transpose) throw std::runtime_error("CudaKernel::spmm_local: Amat needs a non-transposed block");
        if (mode == Bmat && !blk->transpose) throw std::runtime_error("CudaKernel::spmm_local: Bmat needs a transposed block");
        CSRHandle *h = blk->getActive();
        DenseMatrix &In = mode == Amat ? B : A;   // gathered by column index
        DenseMatrix &Out = mode == Amat ? A : B;  // one output row per stored row
        ok(hnh_spmm_host(idx(h->rowStart), idx(h->col_idx), h->values.data(), (int64_t)blk->rows, (int64_t)blk->num_coords,
                         In.data(), (int64_t)In.rows(), Out.data(), (int)A.cols()),
           "hnh_spmm_host");
        return 0;
    }
};
