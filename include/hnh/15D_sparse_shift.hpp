// hnh/15D_sparse_shift.hpp -- 1.5D sparse-shifting algorithm (Sparse15D_Sparse_Shift) on B200s.
//
// Layout and data flow follow the reference (15D_sparse_shift.hpp:23-277, SURVEY.md appendix
// A.2): grid (p/c) x c, adjacency 1; rank (i, j) holds S row block c*i + j (all columns, global
// column indices) as ONE CSR block, and p/c chunks of the dense matrices restricted to R-slice i
// (width R*c/p).  The input dense matrix is all-gathered over row_world chunk by chunk (so each
// rank sees all N rows of its R-slice) and the CSR block -- carrying the partial SDDMM sums in
// its values -- rides a ring over col_world.
//
// What is new here: kernels run directly on row-block VIEWS of the local dense matrix (the
// reference copies each block out to `tmp` and back, :232-249); the CSR shift is a grouped NCCL
// send/recv of (values, col_idx, rowStart) on the communication stream -- for SpMM (block
// read-only) it fully overlaps the kernel, for SDDMM the structure half overlaps and only the
// values wait for the kernel; the declared CSR column count is N, not the reference's too-small
// localArows (:132).
#pragma once
#include <algorithm>

#include "hnh/distributed_sparse.h"

class ShardedBlockRow : public NonzeroDistribution {
public:
    int p, c;
    shared_ptr<FlexibleGrid> grid;
    ShardedBlockRow(int M, int N, int p, int c, shared_ptr<FlexibleGrid> &grid) {
        world = hnh::Comm::world();
        this->p = p;
        this->c = c;
        this->grid = grid;
        rows_in_block = divideAndRoundUp(M, p);
        cols_in_block = N;
    }
    int blockOwner(int row_block, int /*col_block*/) override {
        return grid->get_global_rank(row_block / c, row_block % c, 0);
    }
};

class Sparse15D_Sparse_Shift : public Distributed_Sparse {
public:
    DenseMatrix accumulation_buffer;  // the all-gathered input dense matrix (c > 1)
    int blockAwidth, blockBwidth;
    vector<int> nnz_in_row_axis, nnz_in_row_axis_tpose;

    Sparse15D_Sparse_Shift(SpmatLocal *S_input, int R, int c, KernelImplementation *k) : Distributed_Sparse(k) {
        this->c = c;
        if (c < 1 || p % c != 0) throw hnh::Error(-1, "Error, for 1.5D algorithm, must have c divide num_procs!");
        algorithm_name = "1.5D Sparse Shifting Dense Replicating Algorithm";
        proc_grid_names = {"# Rows", "# Layers"};
        perf_counter_keys = {"Replication Time", "Cyclic Shift Time", "Computation Time"};
        grid.reset(new FlexibleGrid(p / c, c, 1, 1));
        r_split = true;
        A_R_split_world = grid->col_world;
        B_R_split_world = grid->col_world;
        M = (int64_t)S_input->M;
        N = (int64_t)S_input->N;

        ShardedBlockRow standard_dist((int)M, (int)N, p, c, grid);
        ShardedBlockRow transpose_dist((int)N, (int)M, p, c, grid);
        S.reset(S_input->redistribute_nonzeros(&standard_dist, false, false));
        ST.reset(S->redistribute_nonzeros(&transpose_dist, true, false));

        blockAwidth = divideAndRoundUp((int)M, p);
        blockBwidth = divideAndRoundUp((int)N, p);
        localArows = blockAwidth * p / c;
        localBrows = blockBwidth * p / c;
        setRValue(R);

        nnz_in_row_axis = make_monolith(*S, blockAwidth, (int)N);
        nnz_in_row_axis_tpose = make_monolith(*ST, blockBwidth, (int)M);
        check_initialized();
    }

    void setRValue(int R) override {
        this->R = R;
        localAcols = R * c / p;
        localBcols = R * c / p;
        if (localAcols * p / c != R) throw hnh::Error(-1, "Error, R must be divisible by p / c!");
        aSubmatrices.clear();
        bSubmatrices.clear();
        for (int t = 0; t < p / c; t++) {
            aSubmatrices.emplace_back(blockAwidth * (grid->j + c * t), localAcols * grid->i, blockAwidth, localAcols);
            bSubmatrices.emplace_back(blockBwidth * (grid->j + c * t), localBcols * grid->i, blockBwidth, localBcols);
        }
    }

    void initial_shift(DenseMatrix *, DenseMatrix *, KernelMode) override {}
    void de_shift(DenseMatrix *, DenseMatrix *, KernelMode) override {}

    void algorithm(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd *sddmm_result_ptr,
                   KernelMode mode, bool initial_replicate) override {
        hnh::Runtime &rt = hnh::Runtime::get();
        const bool a_mode = (mode == k_spmmA || mode == k_sddmmA);
        const bool sddmm = (mode == k_sddmmA || mode == k_sddmmB);
        DenseMatrix *rowside = a_mode ? &localA : &localB;   // rows of the CSR block index into this
        DenseMatrix *colside = a_mode ? &localB : &localA;   // gathered by column index
        SpmatLocal *choice = a_mode ? S.get() : ST.get();
        const int row_width = a_mode ? blockAwidth : blockBwidth;
        const int col_width = a_mode ? blockBwidth : blockAwidth;
        const vector<int> &nnz_in_axis = a_mode ? nnz_in_row_axis : nnz_in_row_axis_tpose;
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        hnh::Comm &ring = *grid->col_world;
        const int steps = p / c;

        if (initial_replicate && c > 1) gather_colside(*colside, col_width);
        DenseMatrix &gathered = c > 1 ? accumulation_buffer : *colside;

        region_begin("Computation Time", compute());
        if (sddmm) {
            // StandardKernel overwrites at ring step 0, where every block is at home and is
            // visited for the first time; later steps accumulate into the travelling values
            if (!sk) choice->setValuesConstant(0.0);
        } else {
            choice->setCSRValues(SValues);
        }
        region_end("Computation Time", compute());

        CSRLocal *blk = choice->csr_blocks[0];
        const KernelMode local_mode = (mode == k_spmmB) ? k_spmmA : mode;
        const int me = grid->i;
        const int src = pMod(me - 1, steps), dst = pMod(me + 1, steps);
        // copy-engine ring into the next rank's passive CSRHandle when available, NCCL send/recv otherwise
        hnh::PeerRing *pr = steps > 1 ? sparse_ring(grid->col_world, blk) : nullptr;
        for (int t = 0; t < steps; t++) {
            const int block_id = pMod(me - t, steps);
            const int64_t incoming = nnz_in_axis[pMod(me - t - 1, steps)];
            const bool shift = steps > 1;
            const bool early = shift && overlap;
            if (early) {
                // structure (and, for SpMM, the read-only values too) leaves while the kernel runs
                rt.chain(compute(), comm());  // the passive buffer is free: kernel t-1 has been issued before
                region_begin("Cyclic Shift Time", comm());
                if (pr) sparse_push_early(*pr, *blk, sddmm);
                else if (sddmm) blk->shift_structure(src, dst, ring, incoming, comm());
                else blk->shiftCSR_no_flip(src, dst, ring, incoming, comm());
                region_end("Cyclic Shift Time", comm());
            }
            region_begin("Computation Time", compute());
            DenseMatrix rows_view = rowside->rowsView((int64_t)block_id * row_width, row_width);
            if (sk) {
                sk->values_are_zero = sddmm && t == 0;
                sk->output_is_zero = !sddmm;  // each row block of the output is produced exactly once
            } else if (!sddmm) {
                rows_view.setZero();
            }
            kernel->triple_function(local_mode, *choice, rows_view, gathered, 0, 0);
            if (sk) sk->values_are_zero = sk->output_is_zero = false;
            region_end("Computation Time", compute());
            if (shift) {
                rt.chain(compute(), comm());
                region_begin("Cyclic Shift Time", comm());
                if (pr) {
                    sparse_push_late(*pr, *blk, sddmm, early, incoming);
                } else {
                    if (!early) blk->shiftCSR_no_flip(src, dst, ring, incoming, comm());
                    else if (sddmm) blk->shift_values(src, dst, ring, incoming, comm());
                    blk->shift_commit(incoming);
                }
                region_end("Cyclic Shift Time", comm());
                choice->blockStarts[1] = (uint64_t)blk->num_coords;
                rt.chain(comm(), compute());
            }
        }

        if (sddmm) {
            region_begin("Computation Time", compute());
            hadamard_values(*sddmm_result_ptr, SValues, *choice);
            region_end("Computation Time", compute());
        }
    }

    // FusedMM.  When the CSR block never moves (p / c == 1: every rank sees its whole row block of S and, after the
    // all-gather, every row of the other factor at full width) the SDDMM and the SpMM of the reference's two passes
    // (distributed_sparse.h:289-312) are ONE kernel per call: each gathered row is read once, the Hadamard product
    // with Svalues is applied to the dot in registers, sddmm_buffer and the CSR values receive S o dots exactly as the
    // two-pass form leaves them.  Otherwise: the reference's two passes.
    void fusedSpMM(DenseMatrix &localA, DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer, MatMode mode) override {
        StandardKernel *sk = dynamic_cast<StandardKernel *>(kernel);
        const int64_t width = localA.cols();
        const bool table_width = width >= 4 && width <= 256 && (width & (width - 1)) == 0;
        if (sk == nullptr || p / c != 1 || !table_width || width != localB.cols() || sk->sddmm_scale != nullptr) {
            Distributed_Sparse::fusedSpMM(localA, localB, Svalues, sddmm_buffer, mode);
            return;
        }
        const bool a_mode = mode == Amat;
        DenseMatrix &rowside = a_mode ? localA : localB, &colside = a_mode ? localB : localA;
        SpmatLocal *choice = a_mode ? S.get() : ST.get();
        const int64_t total = (int64_t)choice->blockStarts.back();
        if (Svalues.size() != total)
            throw hnh::Error(-1, "SValues has " + to_string(Svalues.size()) + " entries, the local sparse matrix " + to_string(total));
        if (sddmm_buffer.size() != total) sddmm_buffer.resize(total);
        if (c > 1) gather_colside(colside, a_mode ? blockBwidth : blockAwidth);
        DenseMatrix &gathered = c > 1 ? accumulation_buffer : colside;
        region_begin("Computation Time", compute());
        sk->sddmm_scale = &Svalues;
        sk->sddmm_scaled_out = &sddmm_buffer;
        try {
            // in place: output row i depends on input row i of the row-side factor only
            sk->fused_local(*choice, rowside, gathered, rowside, 0, true, true);
        } catch (...) {
            sk->sddmm_scale = nullptr;
            sk->sddmm_scaled_out = nullptr;
            throw;
        }
        sk->sddmm_scale = nullptr;
        sk->sddmm_scaled_out = nullptr;
        region_end("Computation Time", compute());
    }

private:
    // chunk t of every rank of the row world, side by side: all rows of R-slice i in global order (reference :203-215)
    void gather_colside(DenseMatrix &colside, int col_width) {
        hnh::Runtime &rt = hnh::Runtime::get();
        const int64_t cols = colside.cols();
        accumulation_buffer.resize(colside.rows() * c, cols);
        rt.chain(compute(), comm());
        region_begin("Replication Time", comm());
        const size_t chunk = (size_t)col_width * (size_t)cols;
        for (int t = 0; t < p / c; t++)
            grid->row_world->allgather(colside.data() + chunk * t, accumulation_buffer.data() + chunk * c * t,
                                       sizeof(double) * chunk, comm());
        region_end("Replication Time", comm());
        rt.chain(comm(), compute());
    }

    // localise rows, exchange per-rank nnz along the ring, build the single CSR block
    vector<int> make_monolith(SpmatLocal &m, int block_height, int ncols) {
        m.mod_coordinates((uint64_t)block_height, 0);
        vector<int> nnz_in_axis((size_t)(p / c));
        int mine = (int)m.local_tuple_count();
        grid->col_world->host_allgather(&mine, nnz_in_axis.data(), sizeof(int));
        const int max_nnz = *std::max_element(nnz_in_axis.begin(), nnz_in_axis.end());
        m.own_all_coordinates();
        m.monolithBlockColumn();
        m.initializeCSRBlocks(block_height, ncols, max_nnz, false);
        m.release_tuples();
        return nnz_in_axis;
    }
};
