// peer_ring.cpp -- see hnh/peer_ring.h.
#include "hnh/peer_ring.h"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hnh/runtime.h"
#include "hnh_b200.h"

namespace hnh {

namespace {

// driver entry points, fetched through the runtime so that libhnh_b200.so does not link libcuda
// (it must load on a box without a driver)
typedef CUresult (*write32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*wait32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
write32_t g_write32 = nullptr;
wait32_t g_wait32 = nullptr;

void load_driver() {
    if (g_write32 && g_wait32) return;
    cudaDriverEntryPointQueryResult q;
    void *f = nullptr;
    cuda_check(cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q), "cudaGetDriverEntryPoint");
    if (q != cudaDriverEntryPointSuccess || !f) throw Error(HNH_E_CUDA, "cuStreamWriteValue32 not available");
    g_write32 = (write32_t)f;
    cuda_check(cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q), "cudaGetDriverEntryPoint");
    if (q != cudaDriverEntryPointSuccess || !f) throw Error(HNH_E_CUDA, "cuStreamWaitValue32 not available");
    g_wait32 = (wait32_t)f;
}

void cu_check(CUresult r, const char *what) {
    if (r != CUDA_SUCCESS) throw Error(HNH_E_CUDA, std::string(what) + " failed (CUresult " + std::to_string((int)r) + ")");
}

constexpr int kMaxBuffers = 4;
struct Handles {
    cudaIpcMemHandle_t buf[kMaxBuffers][2];
    cudaIpcMemHandle_t flags;
    int nbuf;
    int device;
};

}  // namespace

bool PeerRing::enabled() {
    const char *e = getenv("HNH_RING");
    return !(e && std::strcmp(e, "nccl") == 0);
}

bool PeerRing::all_shifts() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("HNH_RING_ALL");
        v = e ? (atoi(e) != 0 ? 1 : 0) : 1;
    }
    return v == 1 && enabled();
}

PeerRing::PeerRing(std::shared_ptr<Comm> ring, size_t slot_bytes) : ring_(std::move(ring)), bytes_(slot_bytes) {
    if (ring_->size() < 2) throw Error(HNH_E_INVALID, "PeerRing needs at least two ranks");
    // plain cudaMalloc: IPC handles need whole allocations that are never recycled for something else
    owns_slots_ = true;
    bool ok = true;
    for (int k = 0; k < 2 && ok; k++)
        if (cudaMalloc(&slot_[k], bytes_ ? bytes_ : 256) != cudaSuccess) {
            cudaGetLastError();
            slot_[k] = nullptr;
            ok = false;
        }
    {  // collective: a rank that is out of memory must not leave the others waiting in connect()
        int mine = ok ? 1 : 0;
        std::vector<int> all((size_t)ring_->size());
        ring_->host_allgather(&mine, all.data(), sizeof(int));
        for (int v : all)
            if (!v) {
                release_all();
                throw Error(HNH_E_ALLOC, "PeerRing: a rank could not allocate its ring slots");
            }
    }
    connect({{slot_[0], slot_[1]}}, -1);
}

PeerRing::PeerRing(std::shared_ptr<Comm> ring, const std::vector<std::array<void *, 2>> &external, int resident_slot)
    : ring_(std::move(ring)), bytes_(0) {
    if (ring_->size() < 2) throw Error(HNH_E_INVALID, "PeerRing needs at least two ranks");
    if (external.empty() || (int)external.size() > kMaxBuffers) throw Error(HNH_E_INVALID, "PeerRing: 1..4 logical buffers");
    connect(external, resident_slot);
}

// Construction is COLLECTIVE and so is its failure: every step that can fail on one rank only (allocation, IPC
// export, IPC import -- e.g. a topology without uniform peer access) is followed by an exchange of ok-flags over
// the ring, and either every rank goes on or every rank releases what it has acquired and throws.  A rank never
// waits in the final barrier for a neighbour that has already fallen back to NCCL send/recv.
void PeerRing::connect(const std::vector<std::array<void *, 2>> &local, int resident_slot) {
    const int n = ring_->size(), me = ring_->rank();
    const int dst = (me + 1) % n, src = (me + n - 1) % n;
    std::string why;
    auto agree = [&](bool ok_here, const char *stage) {
        int mine = ok_here ? 1 : 0;
        std::vector<int> all((size_t)n);
        ring_->host_allgather(&mine, all.data(), sizeof(int));
        for (int r = 0; r < n; r++)
            if (!all[(size_t)r]) {
                release_all();
                throw Error(HNH_E_COMM, std::string("PeerRing: ") + stage + " failed on ring rank " + std::to_string(r) +
                                            (r == me && !why.empty() ? " (" + why + ")" : std::string()));
            }
    };

    Handles mine;
    std::memset(&mine, 0, sizeof mine);
    bool ok = true;
    try {
        // fault injection for the tests: the named ring rank fails here, as a rank without peer access would
        if (const char *e = getenv("HNH_TEST_PEERRING_FAIL"))
            if (atoi(e) == me) throw Error(HNH_E_CUDA, "injected failure (HNH_TEST_PEERRING_FAIL)");
        load_driver();
        int dev = 0;
        cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
        cuda_check(cudaMalloc((void **)&flags_, 256), "cudaMalloc(ring flags)");
        Flags init;
        std::memset(&init, 0, sizeof init);
        if (resident_slot == 0 || resident_slot == 1) {
            init.arrived[resident_slot] = 1;
            pushed_[resident_slot] = expected_[resident_slot] = 1;
        }
        cuda_check(cudaMemset(flags_, 0, 256), "cudaMemset");
        cuda_check(cudaMemcpy(flags_, &init, sizeof init, cudaMemcpyHostToDevice), "cudaMemcpy(flags)");
        cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
        mine.nbuf = (int)local.size();
        mine.device = dev;
        for (size_t i = 0; i < local.size(); i++)
            for (int k = 0; k < 2; k++) cuda_check(cudaIpcGetMemHandle(&mine.buf[i][k], local[i][(size_t)k]), "cudaIpcGetMemHandle");
        cuda_check(cudaIpcGetMemHandle(&mine.flags, flags_), "cudaIpcGetMemHandle");
    } catch (const Error &e) {
        ok = false;
        why = e.what();
        cudaGetLastError();
    }
    agree(ok, "exporting the ring buffers");

    std::vector<Handles> all((size_t)n);
    ring_->host_allgather(&mine, all.data(), sizeof(Handles));
    try {
        if (all[(size_t)dst].nbuf != mine.nbuf) throw Error(HNH_E_COMM, "ranks registered different buffer counts");
        auto open = [&](const cudaIpcMemHandle_t &h) {
            void *p = nullptr;
            cuda_check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
            return p;
        };
        dst_.assign(local.size(), std::array<void *, 2>{nullptr, nullptr});
        for (size_t i = 0; i < local.size(); i++)
            for (int k = 0; k < 2; k++) dst_[i][(size_t)k] = open(all[(size_t)dst].buf[i][k]);
        dst_flags_ = (Flags *)open(all[(size_t)dst].flags);
        if (src == dst) {
            src_flags_ = dst_flags_;
        } else {
            src_flags_ = (Flags *)open(all[(size_t)src].flags);
            src_flags_opened_ = true;
        }
    } catch (const Error &e) {
        ok = false;
        why = e.what();
        cudaGetLastError();
    }
    agree(ok, "mapping the neighbours' ring buffers");
    ring_->barrier();
}

// Everything connect() / the constructors acquired (also on the failure path: the constructor throws, so the
// destructor does not run).
void PeerRing::release_all() {
    for (auto &b : dst_)
        for (int k = 0; k < 2; k++)
            if (b[(size_t)k]) cudaIpcCloseMemHandle(b[(size_t)k]);
    dst_.clear();
    if (dst_flags_) cudaIpcCloseMemHandle(dst_flags_);
    if (src_flags_opened_ && src_flags_) cudaIpcCloseMemHandle(src_flags_);
    dst_flags_ = src_flags_ = nullptr;
    src_flags_opened_ = false;
    if (owns_slots_) {
        cudaFree(slot_[0]);
        cudaFree(slot_[1]);
        slot_[0] = slot_[1] = nullptr;
        owns_slots_ = false;
    }
    if (flags_) cudaFree(flags_);
    flags_ = nullptr;
    cudaGetLastError();
}

PeerRing::~PeerRing() {
    // make sure nobody is still writing into our memory
    cudaDeviceSynchronize();
    try {
        ring_->barrier();
    } catch (...) {
    }
    release_all();
}

void PeerRing::begin_push(int k, cudaStream_t s) {
    // everything pushed into downstream slot k so far must have been consumed there
    cu_check(g_wait32((CUstream)s, (CUdeviceptr)&flags_->freed[k], pushed_[k], CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
}

void PeerRing::end_push(int k, cudaStream_t s) {
    pushed_[k]++;
    cu_check(g_write32((CUstream)s, (CUdeviceptr)&dst_flags_->arrived[k], pushed_[k], CU_STREAM_WRITE_VALUE_DEFAULT),
             "cuStreamWriteValue32");
}

void PeerRing::push(int k, const void *src, size_t bytes, cudaStream_t s) {
    if (owns_slots_ && bytes > bytes_) throw Error(HNH_E_INVALID, "PeerRing::push: shard larger than the slot");
    begin_push(k, s);
    const int pieces = push_pieces();
    if (pieces <= 1 || bytes < ((size_t)8 << 20)) {
        cuda_check(cudaMemcpyAsync(dst_[0][(size_t)k], src, bytes, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync(peer push)");
    } else {
        // Experiment (HNH_RING_PIECES, off by default): one copy engine may not fill an NVLink port; send the
        // shard as `pieces` concurrent copies, piece 0 on `s`, the others on side streams forked from and
        // joined back into `s` so that the arrival flag is still written after the whole shard.
        static std::vector<cudaStream_t> side;
        static std::vector<cudaEvent_t> done;
        static cudaEvent_t fork = nullptr;
        if (!fork) cuda_check(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming), "cudaEventCreate");
        while ((int)side.size() < pieces - 1) {
            cudaStream_t st;
            cudaEvent_t ev;
            cuda_check(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
            cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
            side.push_back(st);
            done.push_back(ev);
        }
        const size_t piece = ((bytes / (size_t)pieces) + 255) & ~(size_t)255;
        cuda_check(cudaEventRecord(fork, s), "cudaEventRecord");
        for (int i = 0; i < pieces; i++) {
            const size_t off = piece * (size_t)i;
            if (off >= bytes) break;
            const size_t n = std::min(piece, bytes - off);
            cudaStream_t st = i == 0 ? s : side[(size_t)i - 1];
            if (i > 0) cuda_check(cudaStreamWaitEvent(st, fork, 0), "cudaStreamWaitEvent");
            cuda_check(cudaMemcpyAsync((char *)dst_[0][(size_t)k] + off, (const char *)src + off, n, cudaMemcpyDeviceToDevice, st),
                       "cudaMemcpyAsync(peer push piece)");
            if (i > 0) {
                cuda_check(cudaEventRecord(done[(size_t)i - 1], st), "cudaEventRecord");
                cuda_check(cudaStreamWaitEvent(s, done[(size_t)i - 1], 0), "cudaStreamWaitEvent");
            }
        }
    }
    end_push(k, s);
}

int PeerRing::push_pieces() {
    static const int n = [] {
        const char *e = getenv("HNH_RING_PIECES");
        const int v = e ? atoi(e) : 1;
        return v < 1 ? 1 : (v > 8 ? 8 : v);
    }();
    return n;
}

void PeerRing::expect_arrival(int k) { expected_[k]++; }

void PeerRing::wait_arrival(int k, cudaStream_t s) {
    cu_check(g_wait32((CUstream)s, (CUdeviceptr)&flags_->arrived[k], expected_[k], CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
}

void PeerRing::release(int k, cudaStream_t s) {
    consumed_[k]++;
    cu_check(g_write32((CUstream)s, (CUdeviceptr)&src_flags_->freed[k], consumed_[k], CU_STREAM_WRITE_VALUE_DEFAULT),
             "cuStreamWriteValue32");
}

}  // namespace hnh
