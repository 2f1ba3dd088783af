// runtime.cpp -- implementation of hnh::Runtime, EventTimers and the error helpers.
#include "hnh/runtime.h"

#include <chrono>
#include <cstdlib>

#include "hnh_b200.h"

namespace hnh {

void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess)
        throw Error(e == cudaErrorMemoryAllocation ? HNH_E_ALLOC : HNH_E_CUDA,
                    std::string(what) + ": " + cudaGetErrorString(e));
}

void abi_check(int rc, const char *what) {
    if (rc != HNH_OK) throw Error(rc, std::string(what) + ": " + hnh_last_error_string());
}

namespace {
std::map<std::string, double> g_setup_times;
double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace
void setup_time_add(const std::string &phase, double seconds) { g_setup_times[phase] += seconds; }
std::map<std::string, double> setup_times() { return g_setup_times; }
void setup_times_reset() { g_setup_times.clear(); }
SetupPhase::SetupPhase(std::string name) : phase(std::move(name)), t0(wall_now()) {}
SetupPhase::~SetupPhase() { setup_time_add(phase, wall_now() - t0); }
bool device_setup_enabled(int64_t items) {
    static const int forced = [] {
        const char *e = getenv("HNH_DEVICE_SETUP");
        return e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }();
    if (forced == 0 || !Runtime::get().has_device()) return false;
    return forced == 1 || items >= ((int64_t)1 << 18);
}

bool Runtime::managed_mode_ = [] {
    const char *e = getenv("HNH_MANAGED_MEMORY");
    return e != nullptr && atoi(e) != 0;
}();
void host_access_fence_slow() {
    if (Runtime::get().has_device()) Runtime::get().sync_all();
}
int enable_host_access_mode() {
    Runtime::set_managed_mode(true);
    return 1;
}

Runtime &Runtime::get() {
    static Runtime r;
    return r;
}

void Runtime::init() {
    if (inited_) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(HNH_E_CUDA, "no CUDA device: the hnh_b200 host library has no CPU path");
    cuda_check(cudaGetDevice(&dev_), "cudaGetDevice");
    cuda_check(cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking), "cudaStreamCreate");
    cuda_check(cudaStreamCreateWithFlags(&comm_, cudaStreamNonBlocking), "cudaStreamCreate");
    chain_events_.resize(64);
    for (auto &ev : chain_events_)
        cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    inited_ = true;
}

bool Runtime::has_device() {
    if (inited_) return true;
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

cudaStream_t Runtime::compute_stream() { init(); return compute_; }
cudaStream_t Runtime::comm_stream() { init(); return comm_; }
cudaStream_t Runtime::copy_in_stream() {
    init();
    if (!copy_in_) cuda_check(cudaStreamCreateWithFlags(&copy_in_, cudaStreamNonBlocking), "cudaStreamCreate");
    return copy_in_;
}
cudaStream_t Runtime::copy_out_stream() {
    init();
    if (!copy_out_) cuda_check(cudaStreamCreateWithFlags(&copy_out_, cudaStreamNonBlocking), "cudaStreamCreate");
    return copy_out_;
}
int Runtime::device() { init(); return dev_; }

void Runtime::chain(cudaStream_t signaler, cudaStream_t waiter) {
    init();
    if (signaler == waiter) return;
    cudaEvent_t ev = chain_events_[chain_next_++ % chain_events_.size()];
    cuda_check(cudaEventRecord(ev, signaler), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(waiter, ev, 0), "cudaStreamWaitEvent");
}

void Runtime::sync_all() {
    if (!inited_) return;
    cuda_check(cudaStreamSynchronize(compute_), "sync compute");
    cuda_check(cudaStreamSynchronize(comm_), "sync comm");
    if (copy_in_) cuda_check(cudaStreamSynchronize(copy_in_), "sync copy-in");
    if (copy_out_) cuda_check(cudaStreamSynchronize(copy_out_), "sync copy-out");
}

cudaEvent_t Runtime::take_event() {
    if (!free_events_.empty()) {
        cudaEvent_t ev = free_events_.back();
        free_events_.pop_back();
        return ev;
    }
    cudaEvent_t ev;
    cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    return ev;
}

void *Runtime::alloc(size_t bytes) {
    init();
    void *p = nullptr;
    const size_t b = bytes ? ((bytes + 255) & ~(size_t)255) : 256;
    auto it = cache_.find(b);
    if (it != cache_.end() && !it->second.empty()) {
        Cached c = it->second.back();
        it->second.pop_back();
        for (cudaEvent_t ev : c.ev)
            if (ev) {
                cuda_check(cudaStreamWaitEvent(compute_, ev, 0), "cudaStreamWaitEvent");
                free_events_.push_back(ev);
            }
        return c.p;
    }
    auto raw = [&](void **q) { return managed_mode_ ? cudaMallocManaged(q, b, cudaMemAttachGlobal) : cudaMalloc(q, b); };
    cudaError_t e = raw(&p);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
        cudaGetLastError();
        trim();
        e = raw(&p);
    }
    cuda_check(e, managed_mode_ ? "cudaMallocManaged" : "cudaMalloc");
    sizes_[p] = b;
    allocated_ += b;
    return p;
}

void Runtime::free(void *p) {
    if (!p) return;
    auto it = sizes_.find(p);
    if (it == sizes_.end()) {
        cudaFree(p);
        return;
    }
    Cached c{p, {nullptr, nullptr, nullptr}};
    if (inited_) {
        cudaStream_t side[3] = {comm_, copy_in_, copy_out_};
        for (int i = 0; i < 3; i++)
            if (side[i]) {
                cudaEvent_t ev = nullptr;
                try {
                    ev = take_event();
                    cuda_check(cudaEventRecord(ev, side[i]), "cudaEventRecord");
                    c.ev[i] = ev;
                } catch (...) {  // free() runs in destructors: never throw; fall back to draining the stream
                    if (ev) free_events_.push_back(ev);
                    cudaStreamSynchronize(side[i]);
                    cudaGetLastError();
                }
            }
    }
    cache_[it->second].push_back(c);
}

void Runtime::trim() {
    for (auto &kv : cache_)
        for (Cached &c : kv.second) {
            allocated_ -= kv.first;
            sizes_.erase(c.p);
            for (cudaEvent_t ev : c.ev)
                if (ev) free_events_.push_back(ev);
            cudaFree(c.p);
        }
    cache_.clear();
}

void *Runtime::alloc_pinned(size_t bytes) {
    init();
    void *p = nullptr;
    cuda_check(cudaMallocHost(&p, bytes ? bytes : 256), "cudaMallocHost");
    return p;
}

void Runtime::free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- EventTimers ------------
EventTimers::~EventTimers() {
    for (auto &kv : spans_)
        for (auto &s : kv.second) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
    for (auto ev : pool_) cudaEventDestroy(ev);
}

cudaEvent_t EventTimers::get_event() {
    if (!pool_.empty()) {
        cudaEvent_t ev = pool_.back();
        pool_.pop_back();
        return ev;
    }
    cudaEvent_t ev;
    cuda_check(cudaEventCreate(&ev), "cudaEventCreate");
    return ev;
}

void EventTimers::start(const std::string &key, cudaStream_t s) {
    cudaEvent_t ev = get_event();
    cuda_check(cudaEventRecord(ev, s), "cudaEventRecord");
    open_[key] = ev;
}

void EventTimers::stop(const std::string &key, cudaStream_t s) {
    auto it = open_.find(key);
    if (it == open_.end()) throw Error(HNH_E_INVALID, "EventTimers::stop without start: " + key);
    cudaEvent_t b = get_event();
    cuda_check(cudaEventRecord(b, s), "cudaEventRecord");
    spans_[key].push_back({it->second, b});
    counts_[key]++;
    open_.erase(it);
}

void EventTimers::reset() {
    for (auto &kv : spans_)
        for (auto &s : kv.second) {
            pool_.push_back(s.a);
            pool_.push_back(s.b);
        }
    spans_.clear();
    resolved_.clear();
    counts_.clear();
}

double EventTimers::total_seconds(const std::string &key) {
    auto it = spans_.find(key);
    if (it != spans_.end()) {
        double ms = 0.0;
        for (auto &s : it->second) {
            cuda_check(cudaEventSynchronize(s.b), "cudaEventSynchronize");
            float t = 0.f;
            cuda_check(cudaEventElapsedTime(&t, s.a, s.b), "cudaEventElapsedTime");
            ms += t;
            pool_.push_back(s.a);
            pool_.push_back(s.b);
        }
        it->second.clear();
        resolved_[key] += ms * 1e-3;
    }
    return resolved_[key];
}

int EventTimers::count(const std::string &key) const {
    auto it = counts_.find(key);
    return it == counts_.end() ? 0 : it->second;
}

}  // namespace hnh
