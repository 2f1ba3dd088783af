// hnh/distributed_sparse.h -- Distributed_Sparse, the abstract base of the 1.5D / 2.5D
// algorithms (reference distributed_sparse.h:32-388): public operation facade (sddmmA/B,
// spmmA/B, fusedSpMM), layout descriptors, perf counters and the dense ring-shift helper.
//
// B200-native differences:
//   * dense matrices and value vectors live in HBM; every operation is enqueued on
//     hnh::Runtime's streams and returns without waiting for the GPU;
//   * the ring shift is a grouped NCCL send/recv on the communication stream
//     (reference: blocking MPI_Sendrecv from MPI_ANY_SOURCE + a world barrier,
//     distributed_sparse.h:351-361, 15D_dense_shift.hpp:353-356).  NCCL has no ANY_SOURCE: the
//     source is the rank whose destination is this rank (explicit `recv_src`);
//   * `DenseRing` overlaps the transfer of step t+1 with the kernel of step t whenever the
//     riding matrix is only READ by the kernel, using three buffers so the caller's matrix is
//     never overwritten and the last, home-bound shift of the reference is not needed;
//   * perf counters are CUDA-event spans on the stream each region runs on; the reference's
//     keys and JSON schema are kept (distributed_sparse.h:131-179,205-261).
#pragma once
#include <cassert>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "hnh/FlexibleGrid.hpp"
#include "hnh/SpmatLocal.hpp"
#include "hnh/common.h"
#include "hnh/json.h"
#include "hnh/peer_ring.h"
#include "hnh/sparse_kernels.h"

class DenseSubmatrix {
public:
    int topRow, leftCol, rowCount, colCount;
    DenseSubmatrix(int tR, int lC, int rC, int cC) : topRow(tR), leftCol(lC), rowCount(rC), colCount(cC) {}
};

class Distributed_Sparse {
public:
    int proc_rank;  // global rank
    int p, c;       // number of ranks, replication factor

    string algorithm_name;
    vector<string> proc_grid_names;

    vector<string> perf_counter_keys;
    map<string, int> call_count;
    map<string, double> total_time;

    int64_t M, N, R;
    int localArows, localAcols, localBrows, localBcols;
    vector<DenseSubmatrix> aSubmatrices;
    vector<DenseSubmatrix> bSubmatrices;

    unique_ptr<SpmatLocal> S;
    unique_ptr<SpmatLocal> ST;
    shared_ptr<FlexibleGrid> grid;

    int superclass_constructor_sentinel;
    KernelImplementation *kernel;  // non-owning, like the reference

    bool r_split;
    shared_ptr<hnh::Comm> A_R_split_world, B_R_split_world;

    bool verbose;
    string debug_msg;
    // overlap ring shifts with kernels where legal (results never change); on by default
    bool overlap;

    explicit Distributed_Sparse(KernelImplementation *k);
    virtual ~Distributed_Sparse() {}

    virtual void setRValue(int R) = 0;
    void check_initialized();

    json json_algorithm_info();
    void print_algorithm_info();
    void setVerbose(bool value) { verbose = value; }

    // lengths of the value vectors that go with sddmmA/spmmA/fusedSpMM(Amat) (S) and with the
    // B-variants (ST); like_S_values / like_ST_values allocate them on the device
    virtual int64_t num_S_values() { return S->owned_coords_end - S->owned_coords_start; }
    virtual int64_t num_ST_values() { return ST->owned_coords_end - ST->owned_coords_start; }
    virtual VectorXd like_S_values(double value) { return VectorXd::Constant(num_S_values(), value); }
    virtual VectorXd like_ST_values(double value) { return VectorXd::Constant(num_ST_values(), value); }
    DenseMatrix like_A_matrix(double value) { return DenseMatrix::Constant(localArows, localAcols, value); }
    DenseMatrix like_B_matrix(double value) { return DenseMatrix::Constant(localBrows, localBcols, value); }

    // ---- perf counters (keys must be registered in perf_counter_keys) -------------------
    void reset_performance_timers();
    void stop_clock_and_add(my_timer_t &start, string counter_name);  // host wall clock
    void region_begin(const string &counter_name, cudaStream_t s);    // CUDA-event span
    void region_end(const string &counter_name, cudaStream_t s);
    void print_performance_statistics();
    json json_perf_statistics();  // collective; averages over ranks like the reference

    virtual void initial_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode op) = 0;
    virtual void de_shift(DenseMatrix *localA, DenseMatrix *localB, KernelMode op) = 0;

    // ---- the five convenience operations (reference :274-312) ---------------------------
    void spmmA(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues) {
        localA.setZero();
        algorithm(localA, localB, SValues, nullptr, k_spmmA, true);
    }
    void spmmB(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues) {
        localB.setZero();
        algorithm(localA, localB, SValues, nullptr, k_spmmB, true);
    }
    void sddmmA(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd &sddmm_result) {
        algorithm(localA, localB, SValues, &sddmm_result, k_sddmmA, true);
    }
    void sddmmB(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd &sddmm_result) {
        algorithm(localA, localB, SValues, &sddmm_result, k_sddmmB, true);
    }
    // SDDMM then SpMM with "replication reuse": the second pass skips the replication step.
    virtual void fusedSpMM(DenseMatrix &localA, DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer,
                           MatMode mode);

    // fusedSpMM on HOST operands, as the reference's callers hold them (Eigen matrices in host memory):
    // hostA / hostB are this rank's local shards (row-major, like_A_matrix / like_B_matrix shapes), hostOut
    // receives the SpMM result (the shape of the `mode` operand).  localA / localB are the device staging
    // matrices.  Returns when hostOut is complete.  Subclasses may overlap the copies with the kernels.
    virtual void fusedSpMM_host(const double *hostA, const double *hostB, double *hostOut, DenseMatrix &localA,
                                DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer, MatMode mode) {
        localA.copy_from_host(hostA);
        localB.copy_from_host(hostB);
        fusedSpMM(localA, localB, Svalues, sddmm_buffer, mode);
        (mode == Amat ? localA : localB).copy_to_host(hostOut);
    }

    virtual void algorithm(DenseMatrix &localA, DenseMatrix &localB, VectorXd &SValues, VectorXd *sddmm_result_ptr,
                           KernelMode mode, bool initial_replicate) = 0;

    // X[g_row, g_col] = g_row * R + g_col on the local submatrices (reference :322-346)
    void dummyInitialize(DenseMatrix &loc, MatMode mode);

    // Ring-shift a double-buffered matrix: send the active buffer to send_dst, receive the
    // passive one from recv_src, swap.  recv_src < 0: assume a constant-offset ring
    // (src = 2*rank - dst).  Enqueued on stream s (default: the comm stream, chained behind the
    // compute stream and joined back into it).
    void shiftDenseMatrix(BufferPair &buf, hnh::Comm &world, int send_dst, int tag, int recv_src = -1);

    // world-summed squared norms of a local matrix / vector ("fingerprints", scratch.cpp:26-76)
    double fingerprint(const DenseMatrix &m);
    double fingerprint(const VectorXd &v);

protected:
    hnh::EventTimers timers_;
    cudaStream_t compute() { return hnh::Runtime::get().compute_stream(); }
    cudaStream_t comm() { return hnh::Runtime::get().comm_stream(); }

    // result = SValues o (CSR values of all blocks of m), written straight into `result`
    // (reference: `*sddmm_result_ptr = SValues.cwiseProduct(choice->getCSRValues())`, e.g.
    // 15D_dense_shift.hpp:366, which allocates two temporaries and makes three extra passes)
    void hadamard_values(VectorXd &result, VectorXd &SValues, SpmatLocal &m);

    // copy-engine rings (hnh/peer_ring.h), one per (ring communicator, shard size), built on first use
    std::map<std::pair<hnh::Comm *, size_t>, std::unique_ptr<hnh::PeerRing>> peer_rings_;
    bool peer_ring_broken_ = false;
    hnh::PeerRing *peer_ring(std::shared_ptr<hnh::Comm> world, size_t bytes);
    // the same for a riding CSR block: the ring pushes straight into the next rank's passive CSRHandle
    std::map<std::pair<hnh::Comm *, CSRLocal *>, std::unique_ptr<hnh::PeerRing>> sparse_rings_;
    hnh::PeerRing *sparse_ring(std::shared_ptr<hnh::Comm> world, CSRLocal *blk);
    // One ring step of a CSR block over a copy-engine ring.  Call early() before the kernel that uses
    // the block (ships what the kernel does not write), late() after it (ships the rest, publishes,
    // commits).  `values_written`: the kernel accumulates into the block's values (SDDMM).
    void sparse_push_early(hnh::PeerRing &pr, CSRLocal &blk, bool values_written);
    void sparse_push_late(hnh::PeerRing &pr, CSRLocal &blk, bool values_written, bool early_done, int64_t incoming);

    // Driver of a dense ring over `world` (+1 direction, `steps` = world size):
    //   step t: body(t, riding) on the compute stream, then the riding matrix moves one rank on.
    // riding_is_input == true : the kernel only reads the riding matrix.  Three buffers are
    //     used, the caller's matrix is never written, transfers overlap kernels and the last
    //     (home-bound) shift is skipped.
    // riding_is_input == false: the kernel accumulates into the riding matrix (fusion-1 SpMM):
    //     kernel -> shift strictly alternate, `steps` shifts, the result ends up in `home`.
    template <class Body>
    void ring_dense(DenseMatrix &home, std::shared_ptr<hnh::Comm> world_ptr, bool riding_is_input, const string &shift_key,
                    const string &compute_key, Body body);
};

// ------------------------------------------------------------------ ring driver ----------
template <class Body>
void Distributed_Sparse::ring_dense(DenseMatrix &home, std::shared_ptr<hnh::Comm> world_ptr, bool riding_is_input,
                                    const string &shift_key, const string &compute_key, Body body) {
    hnh::Runtime &rt = hnh::Runtime::get();
    hnh::Comm &world = *world_ptr;
    const int steps = world.size();
    const int me = world.rank();
    const int dst = pMod(me + 1, steps), src = pMod(me - 1, steps);
    if (steps == 1) {
        region_begin(compute_key, compute());
        body(0, home);
        region_end(compute_key, compute());
        return;
    }
    const size_t bytes = sizeof(double) * (size_t)home.size();
    hnh::PeerRing *pr = (riding_is_input && overlap) ? peer_ring(world_ptr, bytes) : nullptr;
    if (pr) {
        // Copy engines push the shard of step t into the next rank's slot (t+1)%2 while kernel t
        // runs; flags in peer memory order the two processes (no host, no SM involved).
        rt.chain(compute(), comm());  // `home` is up to date for the first push
        for (int t = 0; t < steps; t++) {
            const int k = t & 1;
            const void *cur = t == 0 ? (const void *)home.data() : pr->slot(k);
            if (t >= 1) pr->expect_arrival(k);
            if (t + 1 < steps) {
                region_begin(shift_key, comm());
                if (t >= 1) pr->wait_arrival(k, comm());
                pr->push((t + 1) & 1, cur, bytes, comm());
                region_end(shift_key, comm());
            }
            if (t >= 1) pr->wait_arrival(k, compute());
            region_begin(compute_key, compute());
            if (t == 0) {
                body(t, home);
            } else {
                DenseMatrix shard = DenseMatrix::view((double *)pr->slot(k), home.rows(), home.cols());
                body(t, shard);
            }
            region_end(compute_key, compute());
            if (t >= 1) {
                rt.chain(compute(), comm());  // slot k is free once kernel t and push t are done
                pr->release(k, comm());
            }
        }
        rt.chain(comm(), compute());
        return;
    }
    if (riding_is_input && overlap) {
        DenseMatrix e1(home.rows(), home.cols()), e2(home.rows(), home.cols());
        DenseMatrix *bufs[3] = {&home, &e1, &e2};
        auto at = [&](int t) -> DenseMatrix * { return t == 0 ? bufs[0] : bufs[1 + ((t - 1) & 1)]; };
        // the comm stream must see everything enqueued so far (home is up to date, e1/e2 free)
        rt.chain(compute(), comm());
        for (int t = 0; t < steps; t++) {
            if (t + 1 < steps) {
                // T_t: at(t) -> next rank's at(t+1).  at(t+1) was last read by K_{t-1}, which the
                // comm stream already waited for (chain below, issued after K_{t-1}).
                region_begin(shift_key, comm());
                world.sendrecv(at(t)->data(), bytes, dst, at(t + 1)->data(), bytes, src, comm());
                region_end(shift_key, comm());
            }
            region_begin(compute_key, compute());
            body(t, *at(t));
            region_end(compute_key, compute());
            if (t + 1 < steps) {
                rt.chain(comm(), compute());  // K_{t+1} needs T_t
                rt.chain(compute(), comm());  // T_{t+1} overwrites at(t+2) == at(t): needs K_t
            }
        }
        rt.chain(comm(), compute());
        return;
    }
    // output rides (or overlap disabled): kernel and shift strictly alternate
    hnh::PeerRing *pro = (!riding_is_input && overlap && hnh::PeerRing::all_shifts()) ? peer_ring(world_ptr, bytes) : nullptr;
    if (pro) {
        // copy engines instead of NCCL p2p (which gets few channels per peer at 8 ranks): the
        // accumulating shard hops through the ring slots and is copied home at the end
        for (int t = 0; t < steps; t++) {
            const int k = t & 1;
            if (t >= 1) {
                pro->expect_arrival(k);
                pro->wait_arrival(k, compute());
            }
            region_begin(compute_key, compute());
            if (t == 0) {
                body(t, home);
            } else {
                DenseMatrix shard = DenseMatrix::view((double *)pro->slot(k), home.rows(), home.cols());
                body(t, shard);
            }
            region_end(compute_key, compute());
            rt.chain(compute(), comm());
            region_begin(shift_key, comm());
            pro->push((t + 1) & 1, t == 0 ? (const void *)home.data() : pro->slot(k), bytes, comm());
            if (t >= 1) pro->release(k, comm());
            region_end(shift_key, comm());
        }
        const int kf = steps & 1;
        pro->expect_arrival(kf);
        region_begin(shift_key, comm());
        pro->wait_arrival(kf, comm());
        hnh::cuda_check(cudaMemcpyAsync(home.data(), pro->slot(kf), bytes, cudaMemcpyDeviceToDevice, comm()), "ring copy-back");
        pro->release(kf, comm());
        region_end(shift_key, comm());
        rt.chain(comm(), compute());
        return;
    }
    BufferPair pair(&home);
    for (int t = 0; t < steps; t++) {
        region_begin(compute_key, compute());
        body(t, *pair.getActive());
        region_end(compute_key, compute());
        rt.chain(compute(), comm());
        region_begin(shift_key, comm());
        world.sendrecv(pair.getActive()->data(), bytes, dst, pair.getPassive()->data(), bytes, src, comm());
        region_end(shift_key, comm());
        pair.swapActive();
        rt.chain(comm(), compute());
    }
    pair.sync_active();
}
