// hnh/25D_cannon_dense.hpp -- 2.5D Cannon's algorithm replicating the dense matrices
// (Sparse25D_Cannon_Dense) on B200s.
//
// Layout and data flow follow the reference (25D_cannon_dense.hpp:26-315, SURVEY.md appendix
// A.3): grid s x s x c with s = sqrt(p/c), adjacency 3; rank (i, j, k) owns dense row block
// k + c*i restricted to R-slice j (width R/s); S is cut into blocks (row block i of height
// ceil(M/(s c))*c, column block j*c + k), stored transposed and pre-skewed along row_world in
// the constructor; per operation: fiber all-gather of the stationary dense operand, then s steps
// of {local kernel; dense shard -> next rank of col_world; sparse block -> next rank of
// row_world}.  The caller brackets operations with initial_shift / de_shift.
//
// What is new here: device-resident operands, NCCL transfers on the communication stream
// overlapped with the kernel wherever the travelling data is only read by it (dense shard and
// CSR structure during SDDMM; the whole CSR block during SpMM).
#pragma once
#include <algorithm>
#include <cmath>

#include "hnh/distributed_sparse.h"

class Block_Cyclic25D : public NonzeroDistribution {
public:
    int sqrtpc, c;
    shared_ptr<FlexibleGrid> grid;
    Block_Cyclic25D(int M, int N, int sqrtpc, int c, shared_ptr<FlexibleGrid> &grid) {
        world = hnh::Comm::world();
        this->sqrtpc = sqrtpc;
        this->c = c;
        this->grid = grid;
        rows_in_block = divideAndRoundUp(M, sqrtpc * c) * c;
        cols_in_block = divideAndRoundUp(N, sqrtpc * c);
    }
    int blockOwner(int row_block, int col_block) override {
        return grid->get_global_rank(row_block, col_block / c, col_block % c);
    }
};

class Sparse25D_Cannon_Dense : public Distributed_Sparse {
public:
    int sqrtpc;
    vector<int> nnz_in_row_axis, nnz_in_row_axis_tpose;
    int sparse_shift;
    DenseMatrix accumulation_buffer;

    Sparse25D_Cannon_Dense(SpmatLocal *S_input, int R, int c, KernelImplementation *k) : Distributed_Sparse(k) {
        this->c = c;
        sqrtpc = (int)std::lround(std::sqrt((double)p / c));
        if (c < 1 || sqrtpc * sqrtpc * c != p)
            throw hnh::Error(-1, "Error, for 2.5D algorithm, p / c must be a perfect square!");
        algorithm_name = "2.5D Cannon's Algorithm Replicating Dense Matrices";
        proc_grid_names = {"# Rows", "# Cols", "# Layers"};
        perf_counter_keys = {"Dense Cyclic Shift Time", "Sparse Cyclic Shift Time", "Dense Fiber Communication Time",
                             "Computation Time", "Setup Shift Time"};
        grid.reset(new FlexibleGrid(sqrtpc, sqrtpc, c, 3));
        A_R_split_world = grid->row_world;
        B_R_split_world = grid->row_world;
        r_split = true;
        M = (int64_t)S_input->M;
        N = (int64_t)S_input->N;
        localArows = divideAndRoundUp((int)M, sqrtpc * c);
        localBrows = divideAndRoundUp((int)N, sqrtpc * c);
        setRValue(R);

        Block_Cyclic25D nonzero_dist((int)M, (int)N, sqrtpc, c, grid);
        Block_Cyclic25D transpose_dist((int)N, (int)M, sqrtpc, c, grid);
        S.reset(S_input->redistribute_nonzeros(&nonzero_dist, false, false));
        ST.reset(S_input->redistribute_nonzeros(&transpose_dist, true, false));

        nnz_in_row_axis = build_block(*S, localArows * c, localBrows);
        nnz_in_row_axis_tpose = build_block(*ST, localBrows * c, localArows);

        // Skew the sparse blocks along the row world for repeated Cannon passes (reference :137-145)
        const int src = pMod(grid->rankInRow + grid->rankInCol, sqrtpc);
        const int dst = pMod(grid->rankInRow - grid->rankInCol, sqrtpc);
        sparse_shift = src;
        S->csr_blocks[0]->shiftCSR(src, dst, *grid->row_world, nnz_in_row_axis[src], 0, both);
        S->blockStarts[1] = (uint64_t)S->csr_blocks[0]->num_coords;
        ST->csr_blocks[0]->shiftCSR(src, dst, *grid->row_world, nnz_in_row_axis_tpose[src], 0, both);
        ST->blockStarts[1] = (uint64_t)ST->csr_blocks[0]->num_coords;
        if (hnh::Runtime::get().has_device()) hnh::Runtime::get().sync_all();
        check_initialized();
    }

    void setRValue(int R) override {
        this->R = R;
        localAcols = R / sqrtpc;
        localBcols = R / sqrtpc;
        if (localAcols * sqrtpc != R) throw hnh::Error(-1, "Error, R must be divisible by sqrt(p / c)!");
        aSubmatrices.clear();
        bSubmatrices.clear();
        aSubmatrices.emplace_back(localArows * (grid->k + c * grid->i), localAcols * grid->j, localArows, localAcols);
        bSubmatrices.emplace_back(localBrows * (grid->k + c * grid->i), localBcols * grid->j, localBrows, localBcols);
    }

    // Align (or restore) the dense operand that rides the ring with the skewed sparse blocks.
    void initial_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode mode) override {
        setup_shift(localA, localB, mode, -1);
    }
    void de_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode mode) override {
        setup_shift(localA, localB, mode, +1);
    }

    // the resident (skewed) block decides the length (reference :214-220)
    int64_t num_S_values() override { return (int64_t)ST->blockStarts[1]; }
    int64_t num_ST_values() override { return (int64_t)S->blockStarts[1]; }

    void algorithm(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd *sddmm_result_ptr,
                   KernelMode mode, bool initial_replicate) override {
        hnh::Runtime &rt = hnh::Runtime::get();
        const bool a_mode = (mode == k_spmmA || mode == k_sddmmA);
        const bool sddmm = (mode == k_sddmmA || mode == k_sddmmB);
        // blocks are stored transposed: for A-modes the stationary operand is B and A rides
        SpmatLocal *choice = a_mode ? ST.get() : S.get();
        DenseMatrix *stationary = a_mode ? &localB : &localA;
        DenseMatrix *riding = a_mode ? &localA : &localB;
        const vector<int> &nnz_in_axis = a_mode ? nnz_in_row_axis_tpose : nnz_in_row_axis;
        if (SValues.size() != (int64_t)choice->blockStarts[1])
            throw hnh::Error(-1, "2.5D dense: SValues length does not match the resident sparse block");
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        const int s = sqrtpc;

        region_begin("Computation Time", compute());
        if (sddmm) {
            if (!sk) choice->setValuesConstant(0.0);
        } else {
            choice->setCSRValues(SValues);
        }
        region_end("Computation Time", compute());

        if (initial_replicate && c > 1) {
            accumulation_buffer.resize(stationary->rows() * c, stationary->cols());
            rt.chain(compute(), comm());
            region_begin("Dense Fiber Communication Time", comm());
            grid->fiber_world->allgather(stationary->data(), accumulation_buffer.data(),
                                         sizeof(double) * (size_t)stationary->size(), comm());
            region_end("Dense Fiber Communication Time", comm());
            rt.chain(comm(), compute());
        }
        DenseMatrix &fixed = c > 1 ? accumulation_buffer : *stationary;

        CSRLocal *blk = choice->csr_blocks[0];
        const size_t dense_bytes = sizeof(double) * (size_t)riding->size();
        const KernelMode local_mode = (mode == k_spmmA) ? k_spmmB : mode;
        hnh::Comm &dense_ring = *grid->col_world, &sparse_ring_comm = *grid->row_world;
        const int d_dst = pMod(grid->rankInCol + 1, s), d_src = pMod(grid->rankInCol - 1, s);
        const int s_dst = pMod(grid->rankInRow + 1, s), s_src = pMod(grid->rankInRow - 1, s);
        auto run_kernel = [&](int t, DenseMatrix &shard) {
            region_begin("Computation Time", compute());
            if (sk) sk->values_are_zero = sddmm && t == 0;
            kernel->triple_function(local_mode, *choice, fixed, shard, 0, localAcols * grid->j);
            if (sk) sk->values_are_zero = false;
            region_end("Computation Time", compute());
        };

        // copy-engine rings (dense shard over col_world, CSR block over row_world) when available
        hnh::PeerRing *prd = (s > 1 && overlap && hnh::PeerRing::all_shifts()) ? peer_ring(grid->col_world, dense_bytes) : nullptr;
        hnh::PeerRing *prs = (prd != nullptr) ? sparse_ring(grid->row_world, blk) : nullptr;
        if (prd && prs) {
            for (int t = 0; t < s; t++) {
                const int64_t incoming = nnz_in_axis[pMod(sparse_shift - t - 1, s)];
                const int kd = (t + 1) & 1;
                const void *cur_ptr = t == 0 ? (const void *)riding->data() : prd->slot(t & 1);
                rt.chain(compute(), comm());
                if (sddmm) {  // the dense shard is an input: it may leave while the kernel reads it
                    region_begin("Dense Cyclic Shift Time", comm());
                    prd->push(kd, cur_ptr, dense_bytes, comm());
                    region_end("Dense Cyclic Shift Time", comm());
                }
                region_begin("Sparse Cyclic Shift Time", comm());
                sparse_push_early(*prs, *blk, sddmm);
                region_end("Sparse Cyclic Shift Time", comm());
                if (t == 0) {
                    run_kernel(t, *riding);
                } else {
                    DenseMatrix shard = DenseMatrix::view((double *)prd->slot(t & 1), riding->rows(), riding->cols());
                    run_kernel(t, shard);
                }
                rt.chain(compute(), comm());
                if (!sddmm) {  // the SpMM output rides: it leaves after the kernel
                    region_begin("Dense Cyclic Shift Time", comm());
                    prd->push(kd, cur_ptr, dense_bytes, comm());
                    region_end("Dense Cyclic Shift Time", comm());
                }
                if (t >= 1) prd->release(t & 1, comm());
                region_begin("Sparse Cyclic Shift Time", comm());
                sparse_push_late(*prs, *blk, sddmm, true, incoming);
                region_end("Sparse Cyclic Shift Time", comm());
                choice->blockStarts[1] = (uint64_t)blk->num_coords;
                prd->expect_arrival(kd);
                prd->wait_arrival(kd, comm());
                rt.chain(comm(), compute());
            }
            // after s hops the dense shard is home again, in slot s & 1
            if (!sddmm)
                hnh::cuda_check(cudaMemcpyAsync(riding->data(), prd->slot(s & 1), dense_bytes, cudaMemcpyDeviceToDevice, comm()),
                                "ring copy-back");
            prd->release(s & 1, comm());
            rt.chain(comm(), compute());
        } else {
        BufferPair ride(riding);
        for (int t = 0; t < s; t++) {
            const int64_t incoming = nnz_in_axis[pMod(sparse_shift - t - 1, s)];
            const bool shift = s > 1;
            const bool early = shift && overlap;
            if (early) {
                rt.chain(compute(), comm());
                if (sddmm) {  // the dense shard is an input: it may leave now
                    region_begin("Dense Cyclic Shift Time", comm());
                    dense_ring.sendrecv(ride.getActive()->data(), dense_bytes, d_dst, ride.getPassive()->data(),
                                        dense_bytes, d_src, comm());
                    region_end("Dense Cyclic Shift Time", comm());
                }
                region_begin("Sparse Cyclic Shift Time", comm());
                if (sddmm) blk->shift_structure(s_src, s_dst, sparse_ring_comm, incoming, comm());
                else blk->shiftCSR_no_flip(s_src, s_dst, sparse_ring_comm, incoming, comm());
                region_end("Sparse Cyclic Shift Time", comm());
            }
            run_kernel(t, *ride.getActive());
            if (shift) {
                rt.chain(compute(), comm());
                if (!early || !sddmm) {  // the SpMM output rides: it leaves after the kernel
                    region_begin("Dense Cyclic Shift Time", comm());
                    dense_ring.sendrecv(ride.getActive()->data(), dense_bytes, d_dst, ride.getPassive()->data(),
                                        dense_bytes, d_src, comm());
                    region_end("Dense Cyclic Shift Time", comm());
                }
                region_begin("Sparse Cyclic Shift Time", comm());
                if (!early) blk->shiftCSR_no_flip(s_src, s_dst, sparse_ring_comm, incoming, comm());
                else if (sddmm) blk->shift_values(s_src, s_dst, sparse_ring_comm, incoming, comm());
                region_end("Sparse Cyclic Shift Time", comm());
                ride.swapActive();
                blk->shift_commit(incoming);
                choice->blockStarts[1] = (uint64_t)blk->num_coords;
                rt.chain(comm(), compute());
            }
        }
        ride.sync_active();
        }

        if (sddmm) {
            region_begin("Computation Time", compute());
            hadamard_values(*sddmm_result_ptr, SValues, *choice);
            region_end("Computation Time", compute());
        }
    }

private:
    vector<int> build_block(SpmatLocal &m, int block_rows, int block_cols) {
        vector<int> nnz_in_axis((size_t)sqrtpc);
        int mine = (int)m.local_tuple_count();
        grid->row_world->host_allgather(&mine, nnz_in_axis.data(), sizeof(int));
        const int max_nnz = *std::max_element(nnz_in_axis.begin(), nnz_in_axis.end());
        m.mod_coordinates((uint64_t)block_rows, (uint64_t)block_cols);
        m.own_all_coordinates();
        m.monolithBlockColumn();
        m.initializeCSRBlocks(block_rows, block_cols, max_nnz, true);
        m.release_tuples();
        return nnz_in_axis;
    }

    // direction -1: initial_shift (send to rankInCol - rankInRow); +1: de_shift
    void setup_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode mode, int direction) {
        DenseMatrix *m = (mode == k_sddmmA || mode == k_spmmA) ? localA : localB;
        if (m == nullptr) return;
        region_begin("Setup Shift Time", compute());
        const int dst = pMod(grid->rankInCol + direction * grid->rankInRow, sqrtpc);
        const int src = pMod(grid->rankInCol - direction * grid->rankInRow, sqrtpc);
        BufferPair buf(m);
        shiftDenseMatrix(buf, *grid->col_world, dst, mode == k_sddmmA || mode == k_spmmA ? 1 : 2, src);
        buf.sync_active();
        region_end("Setup Shift Time", compute());
    }
};
