// als_conjugate_gradients.cpp -- device-resident ALS-CG (see hnh/als_conjugate_gradients.h).
#include "hnh/als_conjugate_gradients.h"

#include <cmath>
#include <iostream>

#include "hnh_b200.h"

using hnh::Runtime;

void ALS_CG::allreduceVector(VectorXd &vec, shared_ptr<hnh::Comm> comm) {
    cudaStream_t s = Runtime::get().compute_stream();
    comm_timer_.start("allreduce", s);
    comm->allreduce_sum_f64(vec.data(), (size_t)vec.size(), s);
    comm_timer_.stop("allreduce", s);
}

void ALS_CG::cg_optimizer(MatMode matrix_to_optimize, int cg_max_iter) {
    const double nan_avoidance_constant = 1e-8;
    shared_ptr<hnh::Comm> reduction_world = matrix_to_optimize == Amat ? A_R_split_world : B_R_split_world;
    DenseMatrix &X = matrix_to_optimize == Amat ? A : B;
    const int64_t nrows = X.rows(), ncols = A.cols();
    const bool r_split = d_ops->r_split;

    DenseMatrix rhs = DenseMatrix::Constant(nrows, ncols, 0.0);
    DenseMatrix Mx(nrows, ncols), Mp(nrows, ncols);
    computeRHS(matrix_to_optimize, rhs);
    computeQueries(A, B, matrix_to_optimize, Mx);

    DenseMatrix r = rhs - Mx;
    DenseMatrix p = r;
    VectorXd rsold = batch_dot_product(r, r);
    VectorXd alpha(rsold.size()), coeffs(rsold.size());
    if (r_split) allreduceVector(rsold, reduction_world);

    for (int cg_iter = 0; cg_iter < cg_max_iter; cg_iter++) {
        if (matrix_to_optimize == Amat) computeQueries(p, B, Amat, Mp);
        else computeQueries(A, p, Bmat, Mp);

        VectorXd bdot = batch_dot_product(p, Mp);
        if (r_split) allreduceVector(bdot, reduction_world);

        // bdot += eps; rsold += eps; alpha = rsold / bdot   (reference .cpp:99-102)
        rsold += nan_avoidance_constant;
        alpha.setQuotient(rsold, 0.0, bdot, nan_avoidance_constant);

        X.setRowAxpy(X, 1.0, &alpha, p);    // X += alpha .* p
        r.setRowAxpy(r, -1.0, &alpha, Mp);  // r -= alpha .* Mp

        VectorXd rsnew = batch_dot_product(r, r);
        if (r_split) allreduceVector(rsnew, reduction_world);

        coeffs.setQuotient(rsnew, 0.0, rsold, 0.0);
        p.setRowAxpy(r, 1.0, &coeffs, p);  // p = r + coeffs .* p
        rsold.swap(rsnew);
    }
    application_communication_time += comm_timer_.total_seconds("allreduce");
    comm_timer_.reset();
}

void ALS_CG::run_cg(int n_alternating_steps) {
    initializeEmbeddings();
    if (proc_rank == 0) std::cout << "Embeddings initialized +" << std::endl;
    for (int i = 0; i < n_alternating_steps; i++) {
        cg_optimizer(Amat, 10);
        cg_optimizer(Bmat, 10);
        if (proc_rank == 0 && i < n_alternating_steps - 1) std::cout << "Completed step " << i << std::endl;
    }
}

static void initialize_dense_matrix(DenseMatrix &X, int64_t R, uint64_t seed) {
    X.setRandom(seed);
    X /= (double)R;
}

Distributed_ALS::Distributed_ALS(Distributed_Sparse *d_ops_in, bool artificial_groundtruth) {
    d_ops = d_ops_in;
    proc_rank = hnh::Comm::world()->rank();
    residual_reduction_world = hnh::Comm::world();
    A_R_split_world = d_ops->A_R_split_world;
    B_R_split_world = d_ops->B_R_split_world;
    application_communication_time = 0.0;
    seed += (uint64_t)proc_rank;

    ones_S_ = d_ops->like_S_values(1.0);
    ones_ST_ = d_ops->like_ST_values(1.0);
    scratch_S_ = d_ops->like_S_values(0.0);
    scratch_ST_ = d_ops->like_ST_values(0.0);

    if (artificial_groundtruth) {
        DenseMatrix Agt = d_ops->like_A_matrix(0.0);
        DenseMatrix Bgt = d_ops->like_B_matrix(0.0);
        initialize_dense_matrix(Agt, d_ops->R, seed * 4 + 0);
        initialize_dense_matrix(Bgt, d_ops->R, seed * 4 + 1);
        Agt /= (double)(d_ops->M * d_ops->R);
        Bgt /= (double)(d_ops->N * d_ops->R);

        // ground truth = SDDMM with all sparse values 1, bracketed by the layout shifts
        ground_truth = d_ops->like_S_values(0.0);
        d_ops->initial_shift(&Agt, &Bgt, k_sddmmA);
        d_ops->sddmmA(Agt, Bgt, ones_S_, ground_truth);
        d_ops->de_shift(&Agt, &Bgt, k_sddmmA);

        ground_truth_transpose = d_ops->like_ST_values(0.0);
        d_ops->initial_shift(&Agt, &Bgt, k_sddmmB);
        d_ops->sddmmB(Agt, Bgt, ones_ST_, ground_truth_transpose);
        d_ops->de_shift(&Agt, &Bgt, k_sddmmB);
    }
}

void Distributed_ALS::computeRHS(MatMode matrix_to_optimize, DenseMatrix &rhs) {
    if (matrix_to_optimize == Amat) {
        d_ops->initial_shift(&rhs, &B, k_spmmA);
        d_ops->spmmA(rhs, B, ground_truth);
        d_ops->de_shift(&rhs, &B, k_spmmA);
    } else {
        d_ops->initial_shift(&A, &rhs, k_spmmB);
        d_ops->spmmB(A, rhs, ground_truth_transpose);
        d_ops->de_shift(&A, &rhs, k_spmmB);
    }
}

double Distributed_ALS::computeResidual() {
    d_ops->initial_shift(&A, &B, k_sddmmA);
    d_ops->sddmmA(A, B, ones_S_, scratch_S_);
    d_ops->de_shift(&A, &B, k_sddmmA);
    double sqnorm = (scratch_S_ - ground_truth).squaredNorm();
    residual_reduction_world->host_allreduce_sum_f64(&sqnorm, 1);
    return std::sqrt(sqnorm);
}

void Distributed_ALS::initializeEmbeddings() {
    A = d_ops->like_A_matrix(1.0);
    B = d_ops->like_B_matrix(1.0);
    initialize_dense_matrix(A, d_ops->R, seed * 4 + 2);
    initialize_dense_matrix(B, d_ops->R, seed * 4 + 3);
    A *= 1.4;
    B /= 1.3;
}

void Distributed_ALS::computeQueries(DenseMatrix &Ain, DenseMatrix &Bin, MatMode matrix_to_optimize, DenseMatrix &result) {
    const double lambda = 1e-13;
    if (matrix_to_optimize == Amat) {
        result = Ain;
        d_ops->initial_shift(&result, &Bin, k_sddmmA);
        d_ops->fusedSpMM(result, Bin, ones_S_, scratch_S_, Amat);
        d_ops->de_shift(&result, &Bin, k_sddmmA);
        result.setRowAxpy(result, lambda, nullptr, Ain);  // result += lambda * A
    } else {
        result = Bin;
        d_ops->initial_shift(&Ain, &result, k_sddmmB);
        d_ops->fusedSpMM(Ain, result, ones_ST_, scratch_ST_, Bmat);
        d_ops->de_shift(&Ain, &result, k_sddmmB);
        result.setRowAxpy(result, lambda, nullptr, Bin);
    }
}
