// hnh/15D_dense_shift.hpp -- 1.5D dense-shifting algorithm (Sparse15D_Dense_Shift) on B200s.
//
// Layout and data flow follow the reference (15D_dense_shift.hpp:22-385, SURVEY.md
// appendix A.1): grid (p/c) x c, adjacency 1; rank (i, j) owns dense row block c*i + j of A and
// of B; S is split by block rows of height ceil(M/p)*c and cyclic block columns (only column
// blocks == j mod c are non-empty on rank (i, j)); the stationary dense operand is replicated
// over row_world (size c) and the other one rides a ring over col_world (size p/c).
// fusionApproach 1 = "replication reuse" (blocks stored transposed, the SpMM OUTPUT rides);
// fusionApproach 2 = "local kernel fusion" (blocks not transposed, inputs ride, the output is
// reduce-scattered) with its own fusedSpMM.
//
// What is new here: device-resident operands; NCCL all-gather / reduce-scatter / ring
// send-recv on a communication stream overlapped with the kernels wherever the riding matrix
// is an input (Distributed_Sparse::ring_dense); one fused SDDMM->SpMM CUDA kernel per ring step
// in the fusion-2 fusedSpMM; persistent replication / accumulation buffers; no per-call
// allocation, zero-fill or copy-back passes that the arithmetic does not need.
#pragma once
#include <memory>

#include "hnh/distributed_sparse.h"

class ShardedBlockCyclicColumn : public NonzeroDistribution {
public:
    int p, c;
    shared_ptr<FlexibleGrid> grid;
    ShardedBlockCyclicColumn(int M, int N, int p, int c, shared_ptr<FlexibleGrid> &grid) {
        world = hnh::Comm::world();
        this->p = p;
        this->c = c;
        this->grid = grid;
        rows_in_block = divideAndRoundUp(M, p) * c;
        cols_in_block = divideAndRoundUp(N, p);
    }
    int blockOwner(int row_block, int col_block) override { return grid->get_global_rank(row_block, col_block % c, 0); }
};

class Sparse15D_Dense_Shift : public Distributed_Sparse {
public:
    int fusionApproach;
    DenseMatrix accumulation_buffer;  // gathered stationary operand, or the SpMM accumulator
    DenseMatrix broadcast_buffer;     // fusion-2 fusedSpMM: gathered stationary operand

    Sparse15D_Dense_Shift(SpmatLocal *S_input, int R, int c, int fusionApproach, KernelImplementation *k)
        : Distributed_Sparse(k) {
        this->fusionApproach = fusionApproach;
        this->c = c;
        if (c < 1 || p % c != 0) throw hnh::Error(-1, "Error, for 1.5D algorithm, must have c divide num_procs!");
        if (fusionApproach != 1 && fusionApproach != 2) throw hnh::Error(-1, "fusionApproach must be 1 or 2");

        algorithm_name = "1.5D Block Row Replicated S Striped AB Cyclic Shift";
        proc_grid_names = {"# Rows", "# Layers"};
        perf_counter_keys = {"Replication Time", "Cyclic Shift Time", "Computation Time"};
        grid.reset(new FlexibleGrid(p / c, c, 1, 1));
        r_split = false;
        M = (int64_t)S_input->M;
        N = (int64_t)S_input->N;

        ShardedBlockCyclicColumn standard_dist((int)M, (int)N, p, c, grid);
        ShardedBlockCyclicColumn transpose_dist((int)N, (int)M, p, c, grid);
        S.reset(S_input->redistribute_nonzeros(&standard_dist, false, false));
        ST.reset(S->redistribute_nonzeros(&transpose_dist, true, false));

        localArows = divideAndRoundUp((int)M, p);
        localBrows = divideAndRoundUp((int)N, p);
        setRValue(R);

        localise(*S, localArows * c, localBrows);
        localise(*ST, localBrows * c, localArows);
        const bool local_tpose = fusionApproach == 1;
        S->initializeCSRBlocks(localArows * c, localBrows, -1, local_tpose);
        S->release_tuples();
        ST->initializeCSRBlocks(localBrows * c, localArows, -1, local_tpose);
        ST->release_tuples();
        check_initialized();
    }

    void setRValue(int R) override {
        this->R = R;
        localAcols = R;
        localBcols = R;
        aSubmatrices.clear();
        bSubmatrices.clear();
        aSubmatrices.emplace_back(localArows * (c * grid->i + grid->j), 0, localArows, localAcols);
        bSubmatrices.emplace_back(localBrows * (c * grid->i + grid->j), 0, localBrows, localBcols);
    }

    void initial_shift(DenseMatrix *, DenseMatrix *, KernelMode) override {}
    void de_shift(DenseMatrix *, DenseMatrix *, KernelMode) override {}

    // fusion 1 keeps the values of an A-mode operation in ST's order (reference :254-270)
    int64_t num_S_values() override {
        SpmatLocal &m = fusionApproach == 1 ? *ST : *S;
        return m.owned_coords_end - m.owned_coords_start;
    }
    int64_t num_ST_values() override {
        SpmatLocal &m = fusionApproach == 1 ? *S : *ST;
        return m.owned_coords_end - m.owned_coords_start;
    }

    // block of S / ST that meets the riding shard at ring step `step`
    int block_at(int step) const { return pMod((grid->rankInCol - step) * c + grid->rankInRow, p); }

    void algorithm(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd *sddmm_result_ptr,
                   KernelMode mode, bool initial_replicate) override {
        const bool a_mode = (mode == k_spmmA || mode == k_sddmmA);
        const bool sddmm = (mode == k_sddmmA || mode == k_sddmmB);
        // fusion 1 works on the transposed problem: for an A-mode operation the stationary
        // (replicated) operand is B and A rides; fusion 2 is the other way round.
        const bool swap_roles = (fusionApproach == 1) == a_mode;
        DenseMatrix *stationary = swap_roles ? &localB : &localA;
        DenseMatrix *riding = swap_roles ? &localA : &localB;
        SpmatLocal *choice = swap_roles ? ST.get() : S.get();
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);

        if (initial_replicate && c > 1) replicate(*stationary, accumulation_buffer);
        DenseMatrix &fixed = c > 1 ? accumulation_buffer : *stationary;

        // StandardKernel: the SDDMM overwrites the block values (every block is met once per pass) and applies the
        // Hadamard product with SValues in its epilogue (reference: a separate pass, :364-368)
        const bool fold = sddmm && sk != nullptr;
        region_begin("Computation Time", compute());
        if (sddmm) {
            if (!sk) choice->setValuesConstant(0.0);
            if (fold) prepare_sddmm_result(*sddmm_result_ptr, SValues, *choice);
        } else if (!spmm_values_resident_) {
            choice->setCSRValues(SValues);
        }
        region_end("Computation Time", compute());

        KernelMode local_mode = mode;
        if (fusionApproach == 1 && mode == k_spmmA) local_mode = k_spmmB;
        if (fusionApproach == 2 && mode == k_spmmB) local_mode = k_spmmA;
        // the riding matrix is written only by the fusion-1 SpMM (it is the output there)
        const bool riding_is_input = sddmm || fusionApproach == 2;

        if (fold) {
            sk->sddmm_scale = &SValues;
            sk->sddmm_scaled_out = sddmm_result_ptr;
            sk->sddmm_scale_values = fused_pass_;  // the SpMM pass of a FusedMM reads S o dots from the CSR values
        }
        ring_dense(*riding, grid->col_world, riding_is_input, "Cyclic Shift Time", "Computation Time",
                   [&](int step, DenseMatrix &shard) {
                       if (sk) {
                           sk->values_are_zero = sddmm;
                           // fusion-1 SpMM: the riding shard is the output; at step 0 it is this rank's own, still
                           // all zeros (the caller's setZero is skipped when the hint is on)
                           sk->output_is_zero = !sddmm && riding_output_is_zero_ && step == 0;
                       }
                       kernel->triple_function(local_mode, *choice, fixed, shard, block_at(step), 0);
                       if (sk) sk->values_are_zero = sk->output_is_zero = false;
                   });
        if (fold) {
            sk->sddmm_scale = nullptr;
            sk->sddmm_scaled_out = nullptr;
            sk->sddmm_scale_values = false;
        }

        if (sddmm && !fold) {
            region_begin("Computation Time", compute());
            hadamard_values(*sddmm_result_ptr, SValues, *choice);
            region_end("Computation Time", compute());
        }
        if (fusionApproach == 2 && !sddmm && c > 1) reduce_to(*stationary, accumulation_buffer);
    }

    // Local kernel fusion: one pass over the ring doing SDDMM and SpMM on each block.
    // Like the reference (15D_dense_shift.hpp:189,250-251) this variant treats S as an all-ones
    // pattern (Svalues ignored) and does not fill sddmm_buffer; the block values it leaves in
    // the CSR are the SDDMM result.
    void fusedSpMM(DenseMatrix &localA, DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer,
                   MatMode mode) override {
        if (fusionApproach == 1) {
            // "replication reuse" (reference distributed_sparse.h:289-312): an SDDMM pass, then an SpMM pass that skips
            // the replication.  With StandardKernel the value plumbing between the two is gone: the SDDMM epilogue
            // leaves Svalues o dots both in sddmm_buffer and in the CSR values (no Hadamard pass, no setCSRValues
            // copy), and the first kernel of the SpMM pass overwrites its output (no zero-fill pass, no read of zeros).
            if (dynamic_cast<StandardKernel *>(kernel) == nullptr) {
                Distributed_Sparse::fusedSpMM(localA, localB, Svalues, sddmm_buffer, mode);
                return;
            }
            fused_pass_ = true;
            try {
                algorithm(localA, localB, Svalues, &sddmm_buffer, mode == Amat ? k_sddmmA : k_sddmmB, true);
                spmm_values_resident_ = riding_output_is_zero_ = true;
                algorithm(localA, localB, sddmm_buffer, nullptr, mode == Amat ? k_spmmA : k_spmmB, false);
            } catch (...) {
                fused_pass_ = spmm_values_resident_ = riding_output_is_zero_ = false;
                throw;
            }
            fused_pass_ = spmm_values_resident_ = riding_output_is_zero_ = false;
            return;
        }
        DenseMatrix *stationary = mode == Amat ? &localA : &localB;
        DenseMatrix *riding = mode == Amat ? &localB : &localA;
        SpmatLocal *choice = mode == Amat ? S.get() : ST.get();
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        const int steps = p / c;

        if (c > 1) replicate(*stationary, broadcast_buffer);
        DenseMatrix &gathered = c > 1 ? broadcast_buffer : *stationary;
        // With one ring step and no replication the output row i depends only on input row i:
        // the fused kernel may then write the result over its own input.
        // (the in-place kernels exist for the widths of the dispatch table only: 4..256, powers of two)
        const bool in_place = sk && steps == 1 && c == 1 && in_place_width(gathered.cols());
        if (!in_place) accumulation_buffer.resize(gathered.rows(), gathered.cols());
        DenseMatrix &out = in_place ? *stationary : accumulation_buffer;
        if (!sk) {
            region_begin("Computation Time", compute());
            choice->setValuesConstant(0.0);
            out.setZero();
            region_end("Computation Time", compute());
        }

        ring_dense(*riding, grid->col_world, true, "Cyclic Shift Time", "Computation Time",
                   [&](int step, DenseMatrix &shard) {
                       kernel->fused_local(*choice, gathered, shard, out, block_at(step), sk != nullptr,
                                           sk != nullptr && step == 0);
                   });

        if (c > 1) {
            reduce_to(*stationary, accumulation_buffer);
        } else if (!in_place) {
            stationary->take(accumulation_buffer);  // instead of `*Arole = accumulation_buffer` (no copy unless a view)
        }
    }

    // Host-operand FusedMM.  With local kernel fusion the result row i depends on row i of the stationary operand
    // and on the riding operand only, so the rows of the stationary operand are streamed: chunk t+1 is uploading
    // and chunk t-1 is downloading (the two PCIe directions) while the fused kernel works on chunk t.  The
    // end-to-end time drops from H2D(A) + H2D(B) + kernels + D2H(out) to about the two uploads plus one chunk.
    // One rank: in place, one kernel per chunk.  Several ranks (c = 1): the riding shards are all-gathered over the
    // ring's communicator while the stationary operand uploads, and every chunk visits all p blocks before it
    // leaves -- the row chunk is the outer loop, the block the inner one.
    int64_t host_pipeline_chunk_rows = 1 << 15;
    void fusedSpMM_host(const double *hostA, const double *hostB, double *hostOut, DenseMatrix &localA,
                        DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer, MatMode mode) override {
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        const int64_t width = (mode == Amat ? localA : localB).cols();
        // (row-range launches need a table width: the generic kernel's overwrite mode clears the values of the
        // whole block, not of a row range)
        if (fusionApproach != 2 || c != 1 || sk == nullptr || host_pipeline_chunk_rows <= 0 || !in_place_width(width)) {
            Distributed_Sparse::fusedSpMM_host(hostA, hostB, hostOut, localA, localB, Svalues, sddmm_buffer, mode);
            return;
        }
        if (p > 1) {
            fusedSpMM_host_gathered(mode == Amat ? hostA : hostB, mode == Amat ? hostB : hostA, hostOut,
                                    mode == Amat ? localA : localB, mode == Amat ? localB : localA,
                                    mode == Amat ? S.get() : ST.get(), sk);
            return;
        }
        hnh::Runtime &rt = hnh::Runtime::get();
        DenseMatrix &stationary = mode == Amat ? localA : localB;
        DenseMatrix &riding = mode == Amat ? localB : localA;
        const double *h_stationary = mode == Amat ? hostA : hostB;
        const double *h_riding = mode == Amat ? hostB : hostA;
        SpmatLocal *choice = mode == Amat ? S.get() : ST.get();
        cudaStream_t in = rt.copy_in_stream(), out = rt.copy_out_stream();
        const int64_t rows = stationary.rows(), r = stationary.cols();

        rt.chain(compute(), in);  // earlier readers / writers of the staging matrices
        if (riding.size())
            hnh::cuda_check(cudaMemcpyAsync(riding.data(), h_riding, sizeof(double) * (size_t)riding.size(),
                                            cudaMemcpyHostToDevice, in), "h2d riding operand");
        for (int64_t row0 = 0; row0 < rows; row0 += host_pipeline_chunk_rows) {
            const int64_t n = std::min(host_pipeline_chunk_rows, rows - row0);
            const size_t bytes = sizeof(double) * (size_t)(n * r);
            hnh::cuda_check(cudaMemcpyAsync(stationary.data() + row0 * r, h_stationary + row0 * r, bytes,
                                            cudaMemcpyHostToDevice, in), "h2d stationary rows");
            rt.chain(in, compute());
            region_begin("Computation Time", compute());
            sk->fused_local_rows(*choice, stationary, riding, stationary, block_at(0), row0, n);
            region_end("Computation Time", compute());
            rt.chain(compute(), out);
            hnh::cuda_check(cudaMemcpyAsync(hostOut + row0 * r, stationary.data() + row0 * r, bytes,
                                            cudaMemcpyDeviceToHost, out), "d2h result rows");
        }
        rt.chain(out, compute());  // later users of the staging matrices are ordered behind the downloads
        hnh::cuda_check(cudaStreamSynchronize(out), "fusedSpMM_host");
    }

private:
    static bool in_place_width(int64_t r) { return r >= 4 && r <= 256 && (r & (r - 1)) == 0; }
    // state of a fusion-1 FusedMM in flight (see fusedSpMM)
    bool fused_pass_ = false, spmm_values_resident_ = false, riding_output_is_zero_ = false;

    // the checks and sizing hadamard_values() does, for the epilogue form of the product
    static void prepare_sddmm_result(VectorXd &result, VectorXd &SValues, SpmatLocal &m) {
        const int64_t total = (int64_t)m.blockStarts.back();
        if (SValues.size() != total)
            throw hnh::Error(-1, "SValues has " + to_string(SValues.size()) + " entries, the local sparse matrix " + to_string(total));
        if (result.size() != total) result.resize(total);
    }

    // p > 1, c = 1: see fusedSpMM_host.  Block b of this rank's block row multiplies the shard owned by rank b of
    // the column communicator (block_at(step) with c = 1 is (rankInCol - step) mod p, the owner of the shard the
    // ring would deliver at that step), so the all-gathered riding operand is indexed by block id.
    void fusedSpMM_host_gathered(const double *h_stationary, const double *h_riding, double *hostOut,
                                 DenseMatrix &stationary, DenseMatrix &riding, SpmatLocal *choice, StandardKernel *sk) {
        hnh::Runtime &rt = hnh::Runtime::get();
        cudaStream_t in = rt.copy_in_stream(), out = rt.copy_out_stream();
        const int64_t rows = stationary.rows(), r = stationary.cols(), shard_rows = riding.rows();

        rt.chain(compute(), in);
        if (riding.size())
            hnh::cuda_check(cudaMemcpyAsync(riding.data(), h_riding, sizeof(double) * (size_t)riding.size(),
                                            cudaMemcpyHostToDevice, in), "h2d riding operand");
        gathered_riding.resize(shard_rows * p, r);
        rt.chain(in, comm());
        rt.chain(compute(), comm());  // earlier readers of the gather buffer
        region_begin("Cyclic Shift Time", comm());
        grid->col_world->allgather(riding.data(), gathered_riding.data(), sizeof(double) * (size_t)riding.size(), comm());
        region_end("Cyclic Shift Time", comm());
        rt.chain(comm(), compute());
        accumulation_buffer.resize(rows, r);

        for (int64_t row0 = 0; row0 < rows; row0 += host_pipeline_chunk_rows) {
            const int64_t n = std::min(host_pipeline_chunk_rows, rows - row0);
            const size_t bytes = sizeof(double) * (size_t)(n * r);
            hnh::cuda_check(cudaMemcpyAsync(stationary.data() + row0 * r, h_stationary + row0 * r, bytes,
                                            cudaMemcpyHostToDevice, in), "h2d stationary rows");
            rt.chain(in, compute());
            region_begin("Computation Time", compute());
            for (int step = 0; step < p; step++) {
                const int b = block_at(step);
                DenseMatrix shard = gathered_riding.rowsView((int64_t)b * shard_rows, shard_rows);
                sk->fused_local_rows(*choice, stationary, shard, accumulation_buffer, b, row0, n, step == 0);
            }
            region_end("Computation Time", compute());
            rt.chain(compute(), out);
            hnh::cuda_check(cudaMemcpyAsync(hostOut + row0 * r, accumulation_buffer.data() + row0 * r, bytes,
                                            cudaMemcpyDeviceToHost, out), "d2h result rows");
        }
        rt.chain(out, compute());
        hnh::cuda_check(cudaStreamSynchronize(out), "fusedSpMM_host");
        stationary.take(accumulation_buffer);  // the staging matrix holds the result, as after fusedSpMM
    }
    DenseMatrix gathered_riding;  // all p shards of the riding operand (fusedSpMM_host, p > 1)

    static void localise(SpmatLocal &m, int block_rows, int block_cols) {
        m.mod_coordinates((uint64_t)block_rows, 0);
        m.divideIntoBlockCols(block_cols, hnh::Comm::world()->size(), true);
        m.own_all_coordinates();
    }

    // all-gather `local` over row_world into `gathered` (rows * c)
    void replicate(DenseMatrix &local, DenseMatrix &gathered) {
        hnh::Runtime &rt = hnh::Runtime::get();
        gathered.resize(local.rows() * c, local.cols());
        rt.chain(compute(), comm());
        region_begin("Replication Time", comm());
        grid->row_world->allgather(local.data(), gathered.data(), sizeof(double) * (size_t)local.size(), comm());
        region_end("Replication Time", comm());
        rt.chain(comm(), compute());
    }

    // reduce-scatter `partial` (rows * c) over row_world into `local`
    void reduce_to(DenseMatrix &local, DenseMatrix &partial) {
        hnh::Runtime &rt = hnh::Runtime::get();
        rt.chain(compute(), comm());
        region_begin("Replication Time", comm());
        grid->row_world->reduce_scatter_sum_f64(partial.data(), local.data(), (size_t)local.size(), comm());
        region_end("Replication Time", comm());
        rt.chain(comm(), compute());
    }
};
