// als_reference_main -- drives the REFERENCE's own als_conjugate_gradients.cpp (compiled UNCHANGED into this
// executable, against include/hnh/compat) on the B200-native classes:
//     als_reference_main logM edgeFactor algorithm R c [cg_iters]
// Ground truth and starting embeddings are a fixed function of the global coordinates (the reference draws them
// with Eigen's setRandom(), which nothing can reproduce), so the result can be compared number for number with the
// library's device-resident ALS on the same inputs (tests/test_dropin.py).  One alternating round:
// cg_optimizer(Amat), cg_optimizer(Bmat) -- the reference's code.  Prints one JSON line on rank 0.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>

#include "als_conjugate_gradients.h"
#include "benchmark_dist.hpp"
#include "sparse_kernels.h"

Distributed_Sparse *hnh_make_algorithm(const string &name, SpmatLocal *spmat, int R, int c, KernelImplementation *k);

// value of global entry (row, col) of operand `salt`: exactly representable, in (-0.5, 0.5) / R
static double entry(uint64_t row, uint64_t col, uint64_t salt, int R) {
    const uint64_t h = (row * 2654435761ull + col * 2246822519ull + salt * 97ull) & 0xffffffffull;
    return ((double)h / 4294967296.0 - 0.5) / (double)R;
}

static void fill(DenseMatrix &m, const vector<DenseSubmatrix> &subs, uint64_t salt, int R, int64_t nrows_global) {
    vector<double> host((size_t)m.size());
    size_t at = 0;
    for (const DenseSubmatrix &s : subs)
        for (int i = 0; i < s.rowCount; i++)
            for (int j = 0; j < s.colCount; j++)
                host[at++] = (s.topRow + i) < nrows_global ? entry((uint64_t)(s.topRow + i), (uint64_t)(s.leftCol + j), salt, R) : 0.0;
    m.copy_from_host(host.data());
}

int main(int argc, char **argv) {
    if (argc < 6) {
        std::cerr << "usage: als_reference_main logM edgeFactor algorithm R c [cg_iters]" << std::endl;
        return 2;
    }
    const int logM = atoi(argv[1]), edgeFactor = atoi(argv[2]);
    const string algorithm_name(argv[3]);
    const int R = atoi(argv[4]), c = atoi(argv[5]), cg_iters = argc > 6 ? atoi(argv[6]) : 10;
    int rc = 0;
    try {
        MPI_Init(&argc, &argv);
        {
            SpmatLocal S;
            S.loadTuples(false, logM, edgeFactor, "");
            StandardKernel kernel;
            std::unique_ptr<Distributed_Sparse> d(hnh_make_algorithm(algorithm_name, &S, R, c, &kernel));
            Distributed_ALS als(d.get(), false);  // the reference's constructor
            als.application_communication_time = 0.0;
            DenseMatrix Agt = d->like_A_matrix(0.0), Bgt = d->like_B_matrix(0.0);
            fill(Agt, d->aSubmatrices, 1, R, d->M);
            fill(Bgt, d->bSubmatrices, 2, R, d->N);
            VectorXd ones = d->like_S_values(1.0);
            als.ground_truth = d->like_S_values(0.0);
            d->initial_shift(&Agt, &Bgt, k_sddmmA);
            d->sddmmA(Agt, Bgt, ones, als.ground_truth);
            d->de_shift(&Agt, &Bgt, k_sddmmA);
            VectorXd ones_t = d->like_ST_values(1.0);
            als.ground_truth_transpose = d->like_ST_values(0.0);
            d->initial_shift(&Agt, &Bgt, k_sddmmB);
            d->sddmmB(Agt, Bgt, ones_t, als.ground_truth_transpose);
            d->de_shift(&Agt, &Bgt, k_sddmmB);
            als.A = d->like_A_matrix(0.0);
            als.B = d->like_B_matrix(0.0);
            fill(als.A, d->aSubmatrices, 3, R, d->M);
            fill(als.B, d->bSubmatrices, 4, R, d->N);
            const double before = als.computeResidual();  // the reference's code from here on
            const double t0 = MPI_Wtime();
            als.cg_optimizer(Amat, cg_iters);
            als.cg_optimizer(Bmat, cg_iters);
            MPI_Barrier(MPI_COMM_WORLD);
            const double seconds = MPI_Wtime() - t0;
            const double after = als.computeResidual();
            const double fa = d->fingerprint(als.A), fb = d->fingerprint(als.B);
            int rank = 0;
            MPI_Comm_rank(MPI_COMM_WORLD, &rank);
            if (rank == 0) {
                std::cout.precision(17);
                std::cout << "{\"algorithm\": \"" << algorithm_name << "\", \"logM\": " << logM << ", \"R\": " << R << ", \"c\": " << c
                          << ", \"cg_iters\": " << cg_iters << ", \"residual_before\": " << before << ", \"residual_after\": " << after
                          << ", \"fingerprint_A\": " << fa << ", \"fingerprint_B\": " << fb << ", \"seconds\": " << seconds
                          << ", \"host_access_mode\": " << (hnh::Runtime::managed_mode() ? "true" : "false") << "}" << std::endl;
            }
        }
        MPI_Finalize();
    } catch (const std::exception &e) {
        std::cerr << "als_reference_main: " << e.what() << std::endl;
        rc = 1;
    }
    return rc;
}
