// hnh/25D_cannon_sparse.hpp -- 2.5D Cannon's algorithm replicating the SPARSE matrix
// (Sparse25D_Cannon_Sparse) on B200s.
//
// Layout and data flow follow the reference (25D_cannon_sparse.hpp:25-314, SURVEY.md appendix
// A.4): grid s x s x c, adjacency 3; the 2-D block (i, j) of S (height ceil(M/s)) is replicated
// on all c layers, layer k owning value segment k (shard_across_layers); the dense matrices are
// split s*c ways along R and BOTH ride rings (classic Cannon: A along row_world, B along
// col_world), s steps; SpMM input values are all-gathered over the fiber, SDDMM partial values
// are reduce-scattered over the fiber.  initial_shift / de_shift exchange a dense matrix with
// the transposed-position rank (j, i, k).
//
// Here the unequal-count fiber collectives (MPI_Allgatherv / MPI_Reduce_scatter of nnz doubles,
// reference :224-233,294-300) are one NCCL all-reduce of the nnz-long value vector each; input
// shards leave for the next rank while the kernel that reads them is still running.
#pragma once
#include <cmath>

#include "hnh/distributed_sparse.h"
#include "hnh_b200.h"

class Floor2D : public NonzeroDistribution {
public:
    shared_ptr<FlexibleGrid> grid;
    Floor2D(int M, int N, int sqrtpc, int /*c*/, shared_ptr<FlexibleGrid> &grid) {
        this->grid = grid;
        world = hnh::Comm::world();
        rows_in_block = divideAndRoundUp(M, sqrtpc);
        cols_in_block = divideAndRoundUp(N, sqrtpc);
    }
    int blockOwner(int row_block, int col_block) override { return grid->get_global_rank(row_block, col_block, 0); }
};

class Sparse25D_Cannon_Sparse : public Distributed_Sparse {
public:
    int sqrtpc;
    int nnz, nnz_tpose;
    VectorXd accumulation_buffer;

    Sparse25D_Cannon_Sparse(SpmatLocal *S_input, int R, int c, KernelImplementation *k) : Distributed_Sparse(k) {
        this->c = c;
        sqrtpc = (int)std::lround(std::sqrt((double)p / c));
        if (c < 1 || sqrtpc * sqrtpc * c != p)
            throw hnh::Error(-1, "Error, for 2.5D algorithm, p / c must be a perfect square!");
        algorithm_name = "2.5D Cannon's Algorithm Replicating Sparse Matrix";
        proc_grid_names = {"# Rows", "# Cols", "# Layers"};
        perf_counter_keys = {"Dense Cyclic Shift Time", "Sparse Fiber Communication Time", "Computation Time",
                             "Setup Shift Time"};
        grid.reset(new FlexibleGrid(sqrtpc, sqrtpc, c, 3));
        A_R_split_world = grid->colfiber_slice();
        B_R_split_world = grid->colfiber_slice();
        r_split = true;
        M = (int64_t)S_input->M;
        N = (int64_t)S_input->N;
        localArows = divideAndRoundUp((int)M, sqrtpc);
        localBrows = divideAndRoundUp((int)N, sqrtpc);
        setRValue(R);

        Floor2D nonzero_dist((int)M, (int)N, sqrtpc, c, grid);
        Floor2D transpose_dist((int)N, (int)M, sqrtpc, c, grid);
        S.reset(S_input->redistribute_nonzeros(&nonzero_dist, false, false));
        ST.reset(S_input->redistribute_nonzeros(&transpose_dist, true, false));
        nnz = replicate_block(*S, localArows, localBrows);
        nnz_tpose = replicate_block(*ST, localBrows, localArows);
        check_initialized();
    }

    void setRValue(int R) override {
        this->R = R;
        localAcols = R / (sqrtpc * c);
        localBcols = R / (sqrtpc * c);
        if (localAcols * sqrtpc * c != R) throw hnh::Error(-1, "Error, R must be divisible by sqrt(pc)!");
        const int shift = pMod(grid->j + grid->i, sqrtpc);
        aSubmatrices.clear();
        bSubmatrices.clear();
        aSubmatrices.emplace_back(localArows * grid->i, localAcols * c * shift + grid->k * localAcols, localArows, localAcols);
        bSubmatrices.emplace_back(localBrows * grid->i, localBcols * c * shift + grid->k * localBcols, localBrows, localBcols);
    }

    // swap the non-output dense matrix with the rank at the transposed grid position
    void initial_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode mode) override {
        DenseMatrix *m = (mode == k_sddmmA || mode == k_spmmA) ? localB : localA;
        if (m == nullptr) return;
        region_begin("Setup Shift Time", compute());
        const int partner = grid->get_global_rank(grid->j, grid->i, grid->k);
        BufferPair buf(m);
        shiftDenseMatrix(buf, *grid->world, partner, 1, partner);
        buf.sync_active();
        region_end("Setup Shift Time", compute());
    }
    void de_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode mode) override {
        initial_shift(localA, localB, mode);
    }

    void algorithm(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd *sddmm_result_ptr,
                   KernelMode mode, bool /*initial_replicate*/) override {
        hnh::Runtime &rt = hnh::Runtime::get();
        const bool a_mode = (mode == k_spmmA || mode == k_sddmmA);
        const bool sddmm = (mode == k_sddmmA || mode == k_sddmmB);
        SpmatLocal *choice = a_mode ? S.get() : ST.get();
        DenseMatrix *rowside = a_mode ? &localA : &localB;  // the SpMM output
        DenseMatrix *colside = a_mode ? &localB : &localA;
        const int64_t nnz_sel = a_mode ? nnz : nnz_tpose;
        const int64_t own0 = choice->owned_coords_start, own_n = choice->owned_coords_end - choice->owned_coords_start;
        if (SValues.size() != own_n) throw hnh::Error(-1, "2.5D sparse: SValues length != owned coordinate count");
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        const int s = sqrtpc;

        if (!sddmm) {
            if (c > 1) {
                // every layer contributes its value segment; the sum over the fiber is the full vector
                region_begin("Sparse Fiber Communication Time", compute());
                accumulation_buffer.resize(nnz_sel);
                accumulation_buffer.setZero();
                if (own_n)
                    hnh::cuda_check(cudaMemcpyAsync(accumulation_buffer.data() + own0, SValues.data(),
                                                    sizeof(double) * (size_t)own_n, cudaMemcpyDeviceToDevice, compute()),
                                    "copy segment");
                grid->fiber_world->allreduce_sum_f64(accumulation_buffer.data(), (size_t)nnz_sel, compute());
                choice->setCSRValues(accumulation_buffer);
                region_end("Sparse Fiber Communication Time", compute());
            } else {
                region_begin("Computation Time", compute());
                choice->setCSRValues(SValues);
                region_end("Computation Time", compute());
            }
        } else if (!sk) {
            region_begin("Computation Time", compute());
            choice->setValuesConstant(0.0);
            region_end("Computation Time", compute());
        }

        const KernelMode local_mode = sddmm ? k_sddmmA : k_spmmA;
        BufferPair rowBuf(rowside), colBuf(colside);
        const size_t row_bytes = sizeof(double) * (size_t)rowside->size(), col_bytes = sizeof(double) * (size_t)colside->size();
        hnh::Comm &row_ring = *grid->row_world, &col_ring = *grid->col_world;
        const int r_dst = pMod(grid->rankInRow + 1, s), r_src = pMod(grid->rankInRow - 1, s);
        const int c_dst = pMod(grid->rankInCol + 1, s), c_src = pMod(grid->rankInCol - 1, s);
        auto send_rows = [&]() {
            row_ring.sendrecv(rowBuf.getActive()->data(), row_bytes, r_dst, rowBuf.getPassive()->data(), row_bytes, r_src, comm());
        };
        auto send_cols = [&]() {
            col_ring.sendrecv(colBuf.getActive()->data(), col_bytes, c_dst, colBuf.getPassive()->data(), col_bytes, c_src, comm());
        };
        for (int t = 0; t < s; t++) {
            const bool shift = s > 1, early = shift && overlap;
            if (early) {  // inputs may leave while the kernel reads them
                rt.chain(compute(), comm());
                region_begin("Dense Cyclic Shift Time", comm());
                if (sddmm) send_rows();
                send_cols();
                region_end("Dense Cyclic Shift Time", comm());
            }
            region_begin("Computation Time", compute());
            if (sk) sk->values_are_zero = sddmm && t == 0;
            kernel->triple_function(local_mode, *choice, *rowBuf.getActive(), *colBuf.getActive(), 0,
                                    pMod(grid->i + grid->j + t, s) * localAcols);
            if (sk) sk->values_are_zero = false;
            region_end("Computation Time", compute());
            if (shift) {
                rt.chain(compute(), comm());
                region_begin("Dense Cyclic Shift Time", comm());
                if (!early || !sddmm) send_rows();  // the SpMM output rides: after the kernel
                if (!early) send_cols();
                region_end("Dense Cyclic Shift Time", comm());
                rowBuf.swapActive();
                colBuf.swapActive();
                rt.chain(comm(), compute());
            }
        }
        rowBuf.sync_active();
        colBuf.sync_active();

        if (sddmm) {
            if (c > 1) {
                region_begin("Sparse Fiber Communication Time", compute());
                choice->getCSRValues(accumulation_buffer);
                grid->fiber_world->allreduce_sum_f64(accumulation_buffer.data(), (size_t)nnz_sel, compute());
                region_end("Sparse Fiber Communication Time", compute());
                region_begin("Computation Time", compute());
                if (sddmm_result_ptr->size() != own_n) sddmm_result_ptr->resize(own_n);
                hnh::abi_check(hnh_hadamard_f64(sddmm_result_ptr->data(), SValues.data(), accumulation_buffer.data() + own0,
                                                own_n, compute()),
                               "hadamard");
                region_end("Computation Time", compute());
            } else {
                region_begin("Computation Time", compute());
                hadamard_values(*sddmm_result_ptr, SValues, *choice);
                region_end("Computation Time", compute());
            }
        }
    }

private:
    // floor layer -> all layers of the fiber; each layer then owns one segment of the values
    int replicate_block(SpmatLocal &m, int block_rows, int block_cols) {
        hnh::Comm &fiber = *grid->fiber_world;
        m.tuples_to_host();  // the fiber broadcast below works on the host tuples
        uint64_t n = m.coords.size();
        vector<uint64_t> counts((size_t)fiber.size());
        fiber.host_allgather(&n, counts.data(), sizeof(uint64_t));
        n = counts[0];
        vector<size_t> sb((size_t)fiber.size(), 0), sd((size_t)fiber.size(), 0), rb((size_t)fiber.size(), 0), rd((size_t)fiber.size(), 0);
        vector<spcoord_t> incoming((size_t)n);
        if (fiber.rank() == 0)
            for (int t = 0; t < fiber.size(); t++) sb[t] = (size_t)n * sizeof(spcoord_t);
        rb[0] = (size_t)n * sizeof(spcoord_t);
        fiber.host_alltoallv(m.coords.data(), sb.data(), sd.data(), incoming.data(), rb.data(), rd.data());
        m.coords.swap(incoming);
        m.shard_across_layers(c, grid->k);
#pragma omp parallel for
        for (int64_t i = 0; i < (int64_t)m.coords.size(); i++) {
            m.coords[i].r %= (uint64_t)block_rows;
            m.coords[i].c %= (uint64_t)block_cols;
        }
        m.monolithBlockColumn();
        m.initializeCSRBlocks(block_rows, block_cols, -1, false);
        const int total = (int)m.coords.size();
        vector<spcoord_t>().swap(m.coords);
        return total;
    }
};
