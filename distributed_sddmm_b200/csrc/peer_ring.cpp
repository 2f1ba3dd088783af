// peer_ring.cpp -- see hnh/peer_ring.h.
#include "hnh/peer_ring.h"

#include <cuda.h>

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

struct Handles {
    cudaIpcMemHandle_t slot[2];
    cudaIpcMemHandle_t flags;
    int device;
    int pid_marker;
};

}  // namespace

bool PeerRing::enabled() {
    const char *e = getenv("HNH_RING");
    return !(e && std::strcmp(e, "nccl") == 0);
}

PeerRing::PeerRing(std::shared_ptr<Comm> ring, size_t slot_bytes) : ring_(std::move(ring)), bytes_(slot_bytes) {
    if (ring_->size() < 2) throw Error(HNH_E_INVALID, "PeerRing needs at least two ranks");
    load_driver();
    int dev = 0;
    cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
    int can = 0;
    cuda_check(cudaDeviceGetAttribute(&can, cudaDevAttrIpcEventSupport, dev), "cudaDeviceGetAttribute");  // proxy for IPC support
    // plain cudaMalloc: IPC handles need whole allocations that are never recycled for something else
    for (int k = 0; k < 2; k++) cuda_check(cudaMalloc(&slot_[k], bytes_ ? bytes_ : 256), "cudaMalloc(ring slot)");
    cuda_check(cudaMalloc((void **)&flags_, 256), "cudaMalloc(ring flags)");
    cuda_check(cudaMemset(flags_, 0, 256), "cudaMemset");
    cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");

    Handles mine;
    std::memset(&mine, 0, sizeof mine);
    for (int k = 0; k < 2; k++) cuda_check(cudaIpcGetMemHandle(&mine.slot[k], slot_[k]), "cudaIpcGetMemHandle");
    cuda_check(cudaIpcGetMemHandle(&mine.flags, flags_), "cudaIpcGetMemHandle");
    mine.device = dev;
    const int n = ring_->size(), me = ring_->rank();
    std::vector<Handles> all((size_t)n);
    ring_->host_allgather(&mine, all.data(), sizeof(Handles));
    const int dst = (me + 1) % n, src = (me + n - 1) % n;
    auto open = [&](const cudaIpcMemHandle_t &h, void **out, int idx) {
        cuda_check(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
        opened_[idx] = true;
    };
    open(all[(size_t)dst].slot[0], &dst_slot_[0], 0);
    open(all[(size_t)dst].slot[1], &dst_slot_[1], 1);
    open(all[(size_t)dst].flags, (void **)&dst_flags_, 2);
    if (src == dst) {
        src_flags_ = dst_flags_;
    } else {
        open(all[(size_t)src].flags, (void **)&src_flags_, 3);
    }
    ring_->barrier();
}

PeerRing::~PeerRing() {
    // make sure nobody is still writing into our memory
    cudaDeviceSynchronize();
    try {
        ring_->barrier();
    } catch (...) {
    }
    if (opened_[0]) cudaIpcCloseMemHandle(dst_slot_[0]);
    if (opened_[1]) cudaIpcCloseMemHandle(dst_slot_[1]);
    if (opened_[2]) cudaIpcCloseMemHandle(dst_flags_);
    if (opened_[3]) cudaIpcCloseMemHandle(src_flags_);
    cudaFree(slot_[0]);
    cudaFree(slot_[1]);
    cudaFree(flags_);
}

void PeerRing::push(int k, const void *src, size_t bytes, cudaStream_t s) {
    if (bytes > bytes_) throw Error(HNH_E_INVALID, "PeerRing::push: shard larger than the slot");
    // everything pushed into downstream slot k so far must have been consumed there
    cu_check(g_wait32((CUstream)s, (CUdeviceptr)&flags_->freed[k], pushed_[k], CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
    cuda_check(cudaMemcpyAsync(dst_slot_[k], src, bytes, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync(peer push)");
    pushed_[k]++;
    cu_check(g_write32((CUstream)s, (CUdeviceptr)&dst_flags_->arrived[k], pushed_[k], CU_STREAM_WRITE_VALUE_DEFAULT),
             "cuStreamWriteValue32");
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
