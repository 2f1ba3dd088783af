// gat_kernels.cu -- the three device operations the GAT forward pass (include/hnh/gat.hpp; reference
// gat.hpp:84-113) needs beyond the SDDMM / SpMM path: LeakyReLU on an edge-value vector, ReLU of a head's
// output written into its column window of the layer output, and the dense projection X * W.
//
// The projection is a plain fp64 library GEMM: cuBLAS (DMMA tensor-core path), bound at first use with
// dlopen so that libhnh_b200.so carries no load-time dependency on libcublas -- everything except
// hnh_dgemm_f64 works without it.
#include <cublas_v2.h>
#include <dlfcn.h>

#include <algorithm>
#include <mutex>

#include "hnh_b200.h"
#include "launch.h"

namespace {

using hnh::check_cuda;
using hnh::count_launch;
using hnh::set_error;

constexpr int kThreads = 256;

// x -> max(x, 0) + alpha * min(x, 0)
__global__ void leaky_relu_kernel(double *__restrict__ dst, const double *__restrict__ src, int64_t n, double alpha) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = src[i];
        dst[i] = fmax(x, 0.0) + fmin(x, 0.0) * alpha;
    }
}

// dst[i, col0 + j] = max(src[i, j], 0): dst has ld_dst columns, src is rows x cols, contiguous
__global__ void relu_cols_kernel(double *__restrict__ dst, int64_t ld_dst, int64_t col0, const double *__restrict__ src,
                                 int64_t rows, int64_t cols) {
    const int64_t n = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const int64_t i = t / cols, j = t - i * cols;
        dst[i * ld_dst + col0 + j] = fmax(src[t], 0.0);
    }
}

int grid_for_elements(int64_t n, int *grid) {
    int sms = 0;
    int rc = hnh::device_sm_count(&sms);
    if (rc) return rc;
    const int64_t want = (n + (int64_t)kThreads * 4 - 1) / ((int64_t)kThreads * 4);
    *grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    return HNH_OK;
}

// ---- cuBLAS, bound lazily -----------------------------------------------------------------------
struct Cublas {
    void *lib = nullptr;
    cublasHandle_t handle = nullptr;
    cublasStatus_t (*create)(cublasHandle_t *) = nullptr;
    cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
    cublasStatus_t (*dgemm)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double *, const double *,
                            int, const double *, int, const double *, double *, int) = nullptr;
    bool failed = false;
};
Cublas g_blas;
std::mutex g_blas_mu;

int bind_cublas() {
    if (g_blas.handle) return HNH_OK;
    if (g_blas.failed) return set_error(HNH_E_INVALID, "hnh_dgemm_f64: cuBLAS is not available in this process");
    g_blas.failed = true;
    for (const char *name : {"libcublas.so.12", "libcublas.so"}) {
        g_blas.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (g_blas.lib) break;
    }
    if (!g_blas.lib) return set_error(HNH_E_INVALID, "hnh_dgemm_f64: cannot load libcublas.so.12 (%s)", dlerror());
    g_blas.create = (decltype(g_blas.create))dlsym(g_blas.lib, "cublasCreate_v2");
    g_blas.set_stream = (decltype(g_blas.set_stream))dlsym(g_blas.lib, "cublasSetStream_v2");
    g_blas.dgemm = (decltype(g_blas.dgemm))dlsym(g_blas.lib, "cublasDgemm_v2");
    if (!g_blas.create || !g_blas.set_stream || !g_blas.dgemm)
        return set_error(HNH_E_INVALID, "hnh_dgemm_f64: libcublas lacks cublasCreate_v2 / cublasSetStream_v2 / cublasDgemm_v2");
    cublasHandle_t h = nullptr;
    const cublasStatus_t st = g_blas.create(&h);
    if (st != CUBLAS_STATUS_SUCCESS) return set_error(HNH_E_CUDA, "hnh_dgemm_f64: cublasCreate failed (status %d)", (int)st);
    g_blas.handle = h;
    g_blas.failed = false;
    return HNH_OK;
}

}  // namespace

extern "C" {

int hnh_leaky_relu_f64(double *dst, const double *src, int64_t n, double alpha, void *stream) {
    if (n < 0) return set_error(HNH_E_INVALID, "hnh_leaky_relu_f64: negative size");
    if (n == 0) return HNH_OK;
    if (!dst || !src) return set_error(HNH_E_INVALID, "hnh_leaky_relu_f64: null pointer");
    int grid;
    int rc = grid_for_elements(n, &grid);
    if (rc) return rc;
    leaky_relu_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(dst, src, n, alpha);
    count_launch(1);
    return check_cuda(cudaGetLastError(), "leaky_relu_kernel launch");
}

int hnh_relu_cols_f64(double *dst, int64_t ld_dst, int64_t col0, const double *src, int64_t rows, int64_t cols, void *stream) {
    if (rows < 0 || cols < 0 || col0 < 0 || ld_dst < col0 + cols)
        return set_error(HNH_E_INVALID, "hnh_relu_cols_f64: column window [%lld, %lld) outside a row of %lld",
                         (long long)col0, (long long)(col0 + cols), (long long)ld_dst);
    if (rows == 0 || cols == 0) return HNH_OK;
    if (!dst || !src) return set_error(HNH_E_INVALID, "hnh_relu_cols_f64: null pointer");
    int grid;
    int rc = grid_for_elements(rows * cols, &grid);
    if (rc) return rc;
    relu_cols_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(dst, ld_dst, col0, src, rows, cols);
    count_launch(1);
    return check_cuda(cudaGetLastError(), "relu_cols_kernel launch");
}

int hnh_dgemm_f64(double *C, const double *A, const double *B, int64_t m, int64_t n, int64_t k, void *stream) {
    if (m < 0 || n < 0 || k < 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX)
        return set_error(HNH_E_INVALID, "hnh_dgemm_f64: bad shape %lld x %lld x %lld", (long long)m, (long long)n, (long long)k);
    if (m == 0 || n == 0) return HNH_OK;
    if (!C || (k > 0 && (!A || !B))) return set_error(HNH_E_INVALID, "hnh_dgemm_f64: null pointer");
    if (k == 0)
        return check_cuda(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)(m * n), (cudaStream_t)stream), "cudaMemsetAsync");
    std::lock_guard<std::mutex> lk(g_blas_mu);
    int rc = bind_cublas();
    if (rc) return rc;
    cublasStatus_t st = g_blas.set_stream(g_blas.handle, (cudaStream_t)stream);
    if (st != CUBLAS_STATUS_SUCCESS) return set_error(HNH_E_CUDA, "hnh_dgemm_f64: cublasSetStream failed (status %d)", (int)st);
    // row-major C (m x n) = A (m x k) B (k x n)  <=>  column-major C^T (n x m) = B^T (n x k) A^T (k x m)
    const double one = 1.0, zero = 0.0;
    st = g_blas.dgemm(g_blas.handle, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)m, (int)k, &one, B, (int)n, A, (int)k, &zero, C, (int)n);
    if (st != CUBLAS_STATUS_SUCCESS) return set_error(HNH_E_CUDA, "hnh_dgemm_f64: cublasDgemm failed (status %d)", (int)st);
    count_launch(1);
    return HNH_OK;
}

}  // extern "C"
