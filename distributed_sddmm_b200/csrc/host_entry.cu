// host_entry.cu -- one-shot host-buffer forms of the two local kernels: everything a reference-side
// `KernelImplementation` needs when the host code (Eigen matrices, MKL_INT vectors) stays as it is
// (INTEGRATION.md route B; include/hnh/reference_plugin/cuda_kernel.h).  Each call mirrors its operands on
// the device, runs the kernel, copies the result back and frees the mirrors: correct for any caller, PCIe-bound
// by construction.  Thread-safe (the reference's ranks may be threads of one process in the test harness).
#include <cuda_runtime.h>

#include "hnh_b200.h"
#include "launch.h"

namespace {

using hnh::check_cuda;
using hnh::set_error;

struct Mirror {  // device copies of one call, released on every exit path
    void *p[8] = {nullptr};
    int n = 0;
    template <class T>
    int up(T **dev, const T *host, size_t count, cudaStream_t st) {
        void *d = nullptr;
        int rc = check_cuda(cudaMalloc(&d, sizeof(T) * (count ? count : 1)), "cudaMalloc");
        if (rc) return rc;
        p[n++] = d;
        *dev = (T *)d;
        if (host && count)
            return check_cuda(cudaMemcpyAsync(d, host, sizeof(T) * count, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync H2D");
        return HNH_OK;
    }
    ~Mirror() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};

}  // namespace

extern "C" {

int hnh_sddmm_coo_host(const int64_t *row_idx, const int64_t *col_idx, double *values, int64_t nnz, const double *X,
                       int64_t x_rows, const double *Y, int64_t y_rows, int r) {
    if (nnz < 0 || x_rows < 0 || y_rows < 0 || r <= 0) return set_error(HNH_E_INVALID, "hnh_sddmm_coo_host: bad argument");
    if (nnz == 0) return HNH_OK;
    if (!row_idx || !col_idx || !values || !X || !Y) return set_error(HNH_E_INVALID, "hnh_sddmm_coo_host: null pointer");
    cudaStream_t st = nullptr;
    Mirror m;
    int64_t *d_ri = nullptr, *d_ci = nullptr;
    double *d_v = nullptr, *d_X = nullptr, *d_Y = nullptr;
    int rc = m.up(&d_ri, row_idx, (size_t)nnz, st);
    if (!rc) rc = m.up(&d_ci, col_idx, (size_t)nnz, st);
    if (!rc) rc = m.up(&d_v, (const double *)values, (size_t)nnz, st);
    if (!rc) rc = m.up(&d_X, X, (size_t)(x_rows * r), st);
    if (!rc) rc = m.up(&d_Y, Y, (size_t)(y_rows * r), st);
    if (!rc) rc = hnh_sddmm_coo_f64(d_ri, d_ci, d_v, nnz, d_X, d_Y, r, 0, st);
    if (!rc) rc = check_cuda(cudaMemcpyAsync(values, d_v, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return rc;
}

int hnh_spmm_host(const int64_t *rowStart, const int64_t *col_idx, const double *values, int64_t rows, int64_t nnz,
                  const double *X, int64_t x_rows, double *Y, int r) {
    if (rows < 0 || nnz < 0 || x_rows < 0 || r <= 0) return set_error(HNH_E_INVALID, "hnh_spmm_host: bad argument");
    if (rows == 0 || nnz == 0) return HNH_OK;
    if (!rowStart || !col_idx || !values || !X || !Y) return set_error(HNH_E_INVALID, "hnh_spmm_host: null pointer");
    cudaStream_t st = nullptr;
    Mirror m;
    int64_t *d_rs = nullptr, *d_ci = nullptr;
    double *d_v = nullptr, *d_X = nullptr, *d_Y = nullptr;
    int rc = m.up(&d_rs, rowStart, (size_t)rows + 1, st);
    if (!rc) rc = m.up(&d_ci, col_idx, (size_t)nnz, st);
    if (!rc) rc = m.up(&d_v, values, (size_t)nnz, st);
    if (!rc) rc = m.up(&d_X, X, (size_t)(x_rows * r), st);
    if (!rc) rc = m.up(&d_Y, (const double *)Y, (size_t)(rows * r), st);
    if (!rc) rc = hnh_spmm_f64(d_rs, d_ci, d_v, rows, nnz, d_X, d_Y, r, 0, st);
    if (!rc) rc = check_cuda(cudaMemcpyAsync(Y, d_Y, sizeof(double) * (size_t)(rows * r), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync D2H");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return rc;
}

}  // extern "C"
