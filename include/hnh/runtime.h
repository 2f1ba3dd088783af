// hnh/runtime.h -- per-process device runtime of the B200-native HnH host library:
// the two CUDA streams every rank works on (compute + communication), stream-ordered
// device allocations, CUDA-event region timers, and the C++ error type.
//
// One process drives one GPU (one rank).  Every device operation of the host classes
// (DenseMatrix algebra, kernels, value plumbing) is enqueued on `compute_stream()`;
// ring shifts / collectives that may overlap with a kernel go on `comm_stream()` and are
// chained with events.  Nothing synchronises with the host unless a result is read back.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace hnh {

// Thrown by the C++ host layer; the extern "C" driver ABI converts it to an HNH_E_* code.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

void cuda_check(cudaError_t e, const char *what);  // throws Error(HNH_E_CUDA)
void abi_check(int rc, const char *what);          // throws Error(rc) with hnh_last_error_string()

class Runtime {
public:
    static Runtime &get();
    // Lazily binds to the current CUDA device and creates the streams.
    cudaStream_t compute_stream();
    cudaStream_t comm_stream();
    // two more streams for host<->device copies that overlap kernels (fusedSpMM_host); created on first use
    cudaStream_t copy_in_stream();
    cudaStream_t copy_out_stream();
    int device();
    bool has_device();
    // make `waiter` wait for everything enqueued on `signaler` so far
    void chain(cudaStream_t signaler, cudaStream_t waiter);
    void sync_all();

    // Host-access (compat) mode: alloc() hands out cudaMallocManaged memory, so that code written against the
    // reference's HOST Eigen matrices -- which takes `double *p = m.data()` and loops over p on the CPU
    // (als_conjugate_gradients.cpp:13-26) -- runs unchanged; DenseMatrix / VectorXd::data() then drain the compute
    // stream before handing the pointer out (host_access_fence).  Off by default: the product keeps everything in
    // plain device memory.  Turned on by HNH_MANAGED_MEMORY=1 or by including the compat <Eigen/Dense>.  Must be chosen
    // before the first allocation.
    static bool managed_mode() { return managed_mode_; }
    static void set_managed_mode(bool on) { managed_mode_ = on; }

    // Caching allocator: free() keeps the block on an exact-size free list and alloc() reuses
    // it, so per-call temporaries (BufferPair::extra, accumulation buffers -- which the
    // reference re-allocates on every algorithm() call, common.h:56-60,
    // 15D_dense_shift.hpp:309) cost no cudaMalloc in steady state.  Reuse is stream-ordered on
    // compute_stream(); free() additionally records an event on every side stream (communication, host copies) and
    // alloc() makes the compute stream wait for them before the block is handed out again, so a block is never
    // recycled under work that was still in flight on another stream when its owner died.
    void *alloc(size_t bytes);  // never returns null; bytes==0 -> 256-byte dummy
    void free(void *p);
    void trim();  // cudaFree everything on the free lists
    void *alloc_pinned(size_t bytes);
    void free_pinned(void *p);
    size_t bytes_allocated() const { return allocated_; }

private:
    Runtime() = default;
    void init();
    static bool managed_mode_;
    bool inited_ = false;
    int dev_ = -1;
    cudaStream_t compute_ = nullptr, comm_ = nullptr, copy_in_ = nullptr, copy_out_ = nullptr;
    std::vector<cudaEvent_t> chain_events_;
    size_t chain_next_ = 0;
    size_t allocated_ = 0;
    std::map<void *, size_t> sizes_;
    // a cached block remembers where the side streams stood when it was freed: whoever reuses it (on the compute
    // stream) first waits for the communication / copy work that had been enqueued by then
    struct Cached {
        void *p;
        cudaEvent_t ev[3];
    };
    std::map<size_t, std::vector<Cached>> cache_;
    std::vector<cudaEvent_t> free_events_;
    cudaEvent_t take_event();
};

// Wall-clock phases of the (untimed) setup path -- tuple generation, redistribution (bucket, exchange, sort),
// COO -> CSR -- accumulated per process; `where` says which side did the heavy lifting ("host" / "device").
// Read through hnhd_setup_times_json(); reset with setup_times_reset().
void setup_time_add(const std::string &phase, double seconds);
std::map<std::string, double> setup_times();
void setup_times_reset();
struct SetupPhase {  // RAII: adds its lifetime to `phase`
    std::string phase;
    double t0;
    explicit SetupPhase(std::string name);
    ~SetupPhase();
};
// HNH_DEVICE_SETUP: "1" force the device-side setup path, "0" force the host path, unset: device when one is
// present and the job is large enough to pay for the copies.
bool device_setup_enabled(int64_t items);

// In host-access mode: everything enqueued on the library's streams so far is complete when this returns (the pointer
// a caller is about to dereference on the CPU is coherent).  A no-op otherwise.
void host_access_fence_slow();
inline void host_access_fence() {
    if (Runtime::managed_mode()) host_access_fence_slow();
}
// for static initialisers of compat headers: switches host-access mode on, returns 1
int enable_host_access_mode();

// Simple owning device array.
template <typename T>
class DeviceBuffer {
public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t n) { resize(n); }
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    DeviceBuffer(DeviceBuffer &&o) noexcept { swap(o); }
    DeviceBuffer &operator=(DeviceBuffer &&o) noexcept {
        if (this != &o) { release(); swap(o); }
        return *this;
    }
    void resize(size_t n) {  // contents are NOT preserved
        if (n <= cap_ && p_) { n_ = n; return; }
        if (!owns_ && p_) throw Error(-1, "DeviceBuffer: cannot grow a non-owning view");
        release();
        p_ = static_cast<T *>(Runtime::get().alloc(n * sizeof(T)));
        n_ = cap_ = n;
    }
    void release() {
        if (p_ && owns_) Runtime::get().free(p_);
        p_ = nullptr; n_ = cap_ = 0; owns_ = true;
    }
    // non-owning window onto memory that belongs to somebody else
    void adopt(T *p, size_t n) {
        release();
        p_ = p; n_ = cap_ = n; owns_ = false;
    }
    bool owns() const { return owns_; }
    void swap(DeviceBuffer &o) noexcept {
        std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); std::swap(owns_, o.owns_);
    }
    T *data() const { return p_; }
    size_t size() const { return n_; }
    void upload(const T *host, size_t n, cudaStream_t s) {
        if (n > n_) resize(n);
        if (n) cuda_check(cudaMemcpyAsync(p_, host, n * sizeof(T), cudaMemcpyHostToDevice, s), "upload");
    }
    std::vector<T> to_host(size_t n, cudaStream_t s) const {
        std::vector<T> h(n);
        if (n) {
            cuda_check(cudaMemcpyAsync(h.data(), p_, n * sizeof(T), cudaMemcpyDeviceToHost, s), "download");
            cuda_check(cudaStreamSynchronize(s), "download sync");
        }
        return h;
    }

private:
    T *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
    bool owns_ = true;
};

// Region timers keyed by name, measured with CUDA events on the stream the region runs on
// (the reference brackets regions with steady_clock, distributed_sparse.h:212-223; on a GPU the
// host clock would only see launch latency).  Elapsed times are resolved lazily in total().
class EventTimers {
public:
    ~EventTimers();
    void start(const std::string &key, cudaStream_t s);
    void stop(const std::string &key, cudaStream_t s);
    void reset();
    double total_seconds(const std::string &key);  // synchronises the recorded events
    int count(const std::string &key) const;

private:
    struct Span { cudaEvent_t a, b; };
    cudaEvent_t get_event();
    std::map<std::string, std::vector<Span>> spans_;
    std::map<std::string, cudaEvent_t> open_;
    std::map<std::string, double> resolved_;
    std::map<std::string, int> counts_;
    std::vector<cudaEvent_t> pool_;
};

}  // namespace hnh
