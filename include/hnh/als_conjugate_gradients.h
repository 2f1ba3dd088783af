// hnh/als_conjugate_gradients.h -- alternating least squares with batched conjugate gradients
// (ALS_CG / Distributed_ALS, reference als_conjugate_gradients.h:23-84 and .cpp:9-301), the
// caller of fusedSpMM in BASELINE.json's config 5, with ALL dense algebra on the device.
//
// The reference runs Eigen expressions and an OpenMP loop on host matrices between every pair
// of fusedSpMM calls (.cpp:9-29,99-138); with the factors living in HBM that would mean PCIe
// round trips of whole shards.  Here batch_dot_product, scale_matrix_rows and the axpy-style
// updates are small bandwidth-bound CUDA kernels (hnh_batch_dot_f64, hnh_row_axpy_f64,
// hnh_vec_quotient_f64) and allreduceVector is an NCCL all-reduce on the R-split communicator.
// The arithmetic (including the 1e-8 "nan avoidance" shifts, lambda = 1e-13 and the absence of
// early stopping) is the reference's.
#pragma once
#include <memory>

#include "hnh/common.h"
#include "hnh/distributed_sparse.h"

class ALS_CG {
public:
    Distributed_Sparse *d_ops;
    DenseMatrix A;
    DenseMatrix B;

    shared_ptr<hnh::Comm> A_R_split_world;
    shared_ptr<hnh::Comm> B_R_split_world;
    shared_ptr<hnh::Comm> residual_reduction_world;

    int proc_rank;
    double application_communication_time;  // seconds of CUDA-event time in allreduceVector

    virtual void computeRHS(MatMode matrix_to_optimize, DenseMatrix &rhs) = 0;
    virtual void computeQueries(DenseMatrix &A, DenseMatrix &B, MatMode matrix_to_optimize, DenseMatrix &result) = 0;
    virtual double computeResidual() = 0;
    virtual void initializeEmbeddings() = 0;

    void allreduceVector(VectorXd &vec, shared_ptr<hnh::Comm> comm);
    void cg_optimizer(MatMode matrix_to_optimize, int cg_max_iter);
    void run_cg(int n_alternating_steps);

    virtual ~ALS_CG() {}

protected:
    hnh::EventTimers comm_timer_;
};

class Distributed_ALS : public ALS_CG {
public:
    VectorXd ground_truth;
    VectorXd ground_truth_transpose;
    uint64_t seed = 1234;  // + rank; Eigen's setRandom() is replaced by a counter-based generator

    Distributed_ALS(Distributed_Sparse *d_ops, bool artificial_groundtruth);

    void computeRHS(MatMode matrix_to_optimize, DenseMatrix &rhs) override;
    void computeQueries(DenseMatrix &A, DenseMatrix &B, MatMode matrix_to_optimize, DenseMatrix &result) override;
    double computeResidual() override;
    void initializeEmbeddings() override;

private:
    // ones / scratch value vectors reused across computeQueries calls (the reference allocates
    // and fills two nnz-long vectors per call, .cpp:278-279,291-292)
    VectorXd ones_S_, ones_ST_, scratch_S_, scratch_ST_;
};
