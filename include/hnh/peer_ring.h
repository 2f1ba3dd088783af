// hnh/peer_ring.h -- copy-engine ring shift over NVLink peer memory.
//
// The dense ring shift of the 1.5D / 2.5D algorithms moves a whole dense shard to the next rank
// at every step (reference: MPI_Sendrecv, distributed_sparse.h:351-361).  With NCCL send/recv the
// copy is done by SM kernels that fight the gather kernel for the memory pipeline: measured on
// B200, a 512 MiB shift takes 0.8 ms alone but 2.2 ms while the fused kernel runs.  The copy
// ENGINES do not have that problem (0.73 ms alone, 0.73 ms under the kernel; profiles/
// r01_ring_contention.md), so the overlapped ring pushes shards with cudaMemcpyAsync into
// buffers of the next rank that are mapped through CUDA IPC, and orders the two processes with
// stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32 on IPC-mapped flag words):
// nothing in the loop touches the host or an SM.
//
// Per rank: two slots (shard-sized) + a few flag words, all visible to the ring neighbours.
//   arrived[k] (my memory, written by my upstream)  = number of shards pushed into my slot k
//   freed[k]   (my memory, written by my downstream) = number of shards it has consumed from ITS slot k
// A push into the downstream slot k waits until everything pushed there before has been consumed;
// a consumer of slot k waits for the arrival it expects.  All counters only grow.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "hnh/comm.h"

namespace hnh {

class PeerRing {
public:
    // Collective over `ring` (size >= 2): allocates the slots, exchanges IPC handles.
    // Throws Error if peer mapping is not possible (the caller then falls back to NCCL).
    PeerRing(std::shared_ptr<Comm> ring, size_t slot_bytes);
    // Same protocol on buffers the caller owns (e.g. the two CSRHandle buffers of a CSRLocal):
    // external[i] = {local pointer of logical buffer i in slot 0, ... in slot 1}; every pointer must be
    // the base of a cudaMalloc allocation (at most 4 logical buffers).  Collective over `ring`.
    // `resident_slot`: the slot that already holds live data on every rank when the ring is built (a
    // CSRLocal's active handle); it is accounted for as if it had been pushed once, so that the first
    // real push into it waits until the downstream rank has consumed its resident block.
    PeerRing(std::shared_ptr<Comm> ring, const std::vector<std::array<void *, 2>> &external, int resident_slot);
    ~PeerRing();
    PeerRing(const PeerRing &) = delete;
    PeerRing &operator=(const PeerRing &) = delete;

    size_t slot_bytes() const { return bytes_; }
    void *slot(int k) const { return slot_[k]; }

    // enqueue on `s`: wait until downstream slot k may be overwritten, copy `bytes` from local
    // `src` into it, then tell the downstream rank that the shard has landed
    void push(int k, const void *src, size_t bytes, cudaStream_t s);
    // the same in pieces: begin_push waits until downstream slot k may be overwritten, the caller then
    // copies into dst_ptr(i, k) with plain cudaMemcpyAsync on the same stream (any number of pieces, on
    // either side of a kernel), end_push publishes the arrival
    void begin_push(int k, cudaStream_t s);
    void *dst_ptr(int i, int k) const { return dst_[(size_t)i][(size_t)k]; }
    void end_push(int k, cudaStream_t s);
    // register that the caller is about to consume the next shard of my slot k (call once per
    // shard), then make stream(s) wait for its arrival
    void expect_arrival(int k);
    void wait_arrival(int k, cudaStream_t s);
    // enqueue on `s`: tell the upstream rank that my slot k has been consumed
    void release(int k, cudaStream_t s);

    static int push_pieces();   // HNH_RING_PIECES (default 1): concurrent copies one push() is split into
    static bool enabled();      // HNH_RING != "nccl"
    static bool all_shifts();   // also use copy-engine rings where the riding data is an output / a CSR block

private:
    struct Flags {
        uint32_t arrived[2];
        uint32_t freed[2];
    };
    std::shared_ptr<Comm> ring_;
    size_t bytes_;
    void *slot_[2] = {nullptr, nullptr};  // owned slots (first constructor only)
    bool owns_slots_ = false;
    Flags *flags_ = nullptr;            // mine
    std::vector<std::array<void *, 2>> dst_;  // downstream's buffers, IPC-mapped
    Flags *dst_flags_ = nullptr;        // downstream's, IPC-mapped
    Flags *src_flags_ = nullptr;        // upstream's, IPC-mapped
    bool src_flags_opened_ = false;
    void connect(const std::vector<std::array<void *, 2>> &local, int resident_slot);
    void release_all();  // idempotent; also the failure path of the (collective) constructors
    uint32_t pushed_[2] = {0, 0};    // shards I pushed into downstream slot k
    uint32_t expected_[2] = {0, 0};  // arrivals into my slot k that have been claimed
    uint32_t consumed_[2] = {0, 0};  // shards of my slot k that I have released
};

}  // namespace hnh
