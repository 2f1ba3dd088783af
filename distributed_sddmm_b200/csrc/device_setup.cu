// device_setup.cu -- device-side versions of the two host setup helpers (host_setup.cpp), exported through
// the C ABI: the seeded Erdos-Renyi tuple generator and the COO -> CSR conversion of one block.  They produce
// bit-identical output to the host versions (the generator is a pure function of (seed, row, k); the CSR
// order is "stored row ascending, input order within a row", i.e. a STABLE sort by stored row, which is what
// cub::DeviceRadixSort gives).  SURVEY.md section 8(f) rank 3: the host classes still build their blocks
// with the host path; these entry points are the building blocks of a device-resident setup.
#include <cub/cub.cuh>

#include <algorithm>

#include "hnh_b200.h"
#include "launch.h"

namespace {

using hnh::check_cuda;
using hnh::count_launch;
using hnh::set_error;

constexpr int kThreads = 256;
constexpr int kMaxPerRow = 128;  // columns drawn per row that one thread sorts in local memory

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// the sorted, de-duplicated column draws of one row; returns how many are left
__device__ inline int draw_row(uint64_t seed, uint64_t row, int per_row, uint64_t mask, uint64_t *tmp) {
    const uint64_t base = mix64(seed ^ (row * 0xD1342543DE82EF95ull));
    for (int k = 0; k < per_row; k++) {  // insertion sort while drawing
        const uint64_t x = mix64(base + (uint64_t)k) & mask;
        int j = k;
        while (j > 0 && tmp[j - 1] > x) { tmp[j] = tmp[j - 1]; j--; }
        tmp[j] = x;
    }
    int n = 0;
    for (int k = 0; k < per_row; k++)
        if (k == 0 || tmp[k] != tmp[k - 1]) tmp[n++] = tmp[k];
    return n;
}

__global__ void er_count_kernel(uint64_t seed, int64_t row_lo, int64_t nrows, int per_row, uint64_t mask, int64_t *counts) {
    uint64_t tmp[kMaxPerRow];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride)
        counts[i] = draw_row(seed, (uint64_t)(row_lo + i), per_row, mask, tmp);
}

__global__ void er_write_kernel(uint64_t seed, int64_t row_lo, int64_t nrows, int per_row, uint64_t mask,
                                const int64_t *__restrict__ offsets, uint64_t *__restrict__ rows_out,
                                uint64_t *__restrict__ cols_out, double *__restrict__ vals_out) {
    uint64_t tmp[kMaxPerRow];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += stride) {
        const int n = draw_row(seed, (uint64_t)(row_lo + i), per_row, mask, tmp);
        int64_t o = offsets[i];
        for (int k = 0; k < n; k++, o++) {
            rows_out[o] = (uint64_t)(row_lo + i);
            cols_out[o] = tmp[k];
            vals_out[o] = 1.0;
        }
    }
}

// key[i] = stored row of entry i, idx[i] = i; *bad is set when a coordinate lies outside the block
__global__ void csr_keys_kernel(const uint64_t *__restrict__ sr, const uint64_t *__restrict__ sc, int64_t nnz,
                                uint64_t out_rows, uint64_t out_cols, uint64_t *__restrict__ key, int64_t *__restrict__ idx,
                                int *bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        if (sr[i] >= out_rows || sc[i] >= out_cols) *bad = 1;
        key[i] = sr[i];
        idx[i] = i;
    }
}

__global__ void csr_gather_kernel(const uint64_t *__restrict__ sorted_key, const int64_t *__restrict__ sorted_idx,
                                  const uint64_t *__restrict__ sc, const double *__restrict__ v, int64_t nnz,
                                  int64_t *__restrict__ col_idx, int64_t *__restrict__ row_idx, double *__restrict__ values) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) {
        const int64_t i = sorted_idx[p];
        col_idx[p] = (int64_t)sc[i];
        values[p] = v[i];
        if (row_idx) row_idx[p] = (int64_t)sorted_key[p];
    }
}

// rowStart[r] = first position p with sorted_key[p] >= r, for r = 0 .. out_rows (binary search per row)
__global__ void csr_row_start_kernel(const uint64_t *__restrict__ sorted_key, int64_t nnz, int64_t out_rows,
                                     int64_t *__restrict__ rowStart) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= out_rows; r += stride) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (sorted_key[mid] < (uint64_t)r) lo = mid + 1; else hi = mid;
        }
        rowStart[r] = lo;
    }
}

int grid_for(int64_t n) {
    int sms = 148;
    hnh::device_sm_count(&sms);
    const int64_t want = (n + kThreads - 1) / kThreads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
}

struct Scratch {  // cudaMalloc'ed temporaries of one call, released on every exit path
    void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    cudaError_t get(void **out, size_t bytes) {
        cudaError_t e = cudaMalloc(out, std::max<size_t>(bytes, 256));
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
    ~Scratch() { for (int i = 0; i < n; i++) cudaFree(p[i]); }
};

int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) b++;
    return b;
}

}  // namespace

extern "C" {

int64_t hnh_er_generate_device(int logM, int nnz_per_row, uint64_t seed, int64_t row_lo, int64_t row_hi, uint64_t *rows_out,
                               uint64_t *cols_out, double *vals_out, int64_t capacity, void *stream) {
    if (logM < 0 || logM > 40 || nnz_per_row < 0 || nnz_per_row > kMaxPerRow || row_lo < 0 || row_hi < row_lo ||
        row_hi > ((int64_t)1 << logM))
        return set_error(HNH_E_INVALID, "hnh_er_generate_device: bad argument (nnz_per_row <= %d)", kMaxPerRow);
    const int64_t nrows = row_hi - row_lo;
    if (nrows == 0 || nnz_per_row == 0) return 0;
    if (nrows >= INT32_MAX) return set_error(HNH_E_INVALID, "hnh_er_generate_device: at most 2^31 - 2 rows per call");
    if (!rows_out || !cols_out || !vals_out) return set_error(HNH_E_INVALID, "hnh_er_generate_device: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t mask = ((uint64_t)1 << logM) - 1;
    Scratch s;
    int64_t *counts = nullptr, *offsets = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = check_cuda(s.get((void **)&counts, sizeof(int64_t) * (size_t)(nrows + 1)), "cudaMalloc");
    if (!rc) rc = check_cuda(s.get((void **)&offsets, sizeof(int64_t) * (size_t)(nrows + 1)), "cudaMalloc");
    if (rc) return rc;
    er_count_kernel<<<grid_for(nrows), kThreads, 0, st>>>(seed, row_lo, nrows, nnz_per_row, mask, counts);
    count_launch(1);
    rc = check_cuda(cudaMemsetAsync(counts + nrows, 0, sizeof(int64_t), st), "cudaMemsetAsync");
    if (rc) return rc;
    rc = check_cuda(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, offsets, (int)(nrows + 1), st), "cub scan size");
    if (!rc) rc = check_cuda(s.get(&tmp, tmp_bytes), "cudaMalloc");
    if (!rc) rc = check_cuda(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, offsets, (int)(nrows + 1), st), "cub scan");
    if (rc) return rc;
    int64_t total = 0;
    rc = check_cuda(cudaMemcpyAsync(&total, offsets + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    if (rc) return rc;
    if (total > capacity)
        return set_error(HNH_E_INVALID, "hnh_er_generate_device: capacity %lld < %lld", (long long)capacity, (long long)total);
    er_write_kernel<<<grid_for(nrows), kThreads, 0, st>>>(seed, row_lo, nrows, nnz_per_row, mask, offsets, rows_out, cols_out, vals_out);
    count_launch(1);
    rc = check_cuda(cudaGetLastError(), "er_write_kernel launch");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");  // the scratch is freed on return
    return rc ? rc : total;
}

int hnh_coo_to_csr_device(int64_t rows, int64_t cols, int64_t nnz, const uint64_t *r, const uint64_t *c, const double *v,
                          int transpose, int64_t *rowStart, int64_t *col_idx, int64_t *row_idx, double *values, void *stream) {
    if (rows < 0 || cols < 0 || nnz < 0 || !rowStart) return set_error(HNH_E_INVALID, "hnh_coo_to_csr_device: bad argument");
    if (nnz > 0 && (!r || !c || !v || !col_idx || !values)) return set_error(HNH_E_INVALID, "hnh_coo_to_csr_device: null pointer");
    if (nnz > INT32_MAX) return set_error(HNH_E_INVALID, "hnh_coo_to_csr_device: more than 2^31 - 1 entries in one block");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t out_rows = transpose ? cols : rows, out_cols = transpose ? rows : cols;
    const uint64_t *sr = transpose ? c : r, *sc = transpose ? r : c;
    if (nnz == 0) return check_cuda(cudaMemsetAsync(rowStart, 0, sizeof(int64_t) * (size_t)(out_rows + 1), st), "cudaMemsetAsync");
    Scratch s;
    uint64_t *key = nullptr, *key_sorted = nullptr;
    int64_t *idx = nullptr, *idx_sorted = nullptr;
    int *bad = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int rc = check_cuda(s.get((void **)&key, sizeof(uint64_t) * (size_t)nnz), "cudaMalloc");
    if (!rc) rc = check_cuda(s.get((void **)&key_sorted, sizeof(uint64_t) * (size_t)nnz), "cudaMalloc");
    if (!rc) rc = check_cuda(s.get((void **)&idx, sizeof(int64_t) * (size_t)nnz), "cudaMalloc");
    if (!rc) rc = check_cuda(s.get((void **)&idx_sorted, sizeof(int64_t) * (size_t)nnz), "cudaMalloc");
    if (!rc) rc = check_cuda(s.get((void **)&bad, sizeof(int)), "cudaMalloc");
    if (!rc) rc = check_cuda(cudaMemsetAsync(bad, 0, sizeof(int), st), "cudaMemsetAsync");
    if (rc) return rc;
    csr_keys_kernel<<<grid_for(nnz), kThreads, 0, st>>>(sr, sc, nnz, (uint64_t)out_rows, (uint64_t)out_cols, key, idx, bad);
    count_launch(1);
    int host_bad = 0;
    rc = check_cuda(cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    if (rc) return rc;
    if (host_bad)
        return set_error(HNH_E_INVALID, "hnh_coo_to_csr_device: a coordinate lies outside the %lld x %lld block",
                         (long long)rows, (long long)cols);
    const int end_bit = bits_for((uint64_t)std::max<int64_t>(out_rows - 1, 1));
    rc = check_cuda(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key_sorted, idx, idx_sorted, (int)nnz, 0, end_bit, st),
                    "cub sort size");
    if (!rc) rc = check_cuda(s.get(&tmp, tmp_bytes), "cudaMalloc");
    if (!rc) rc = check_cuda(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key_sorted, idx, idx_sorted, (int)nnz, 0, end_bit, st),
                             "cub sort");
    if (rc) return rc;
    csr_gather_kernel<<<grid_for(nnz), kThreads, 0, st>>>(key_sorted, idx_sorted, sc, v, nnz, col_idx, row_idx, values);
    csr_row_start_kernel<<<grid_for(out_rows + 1), kThreads, 0, st>>>(key_sorted, nnz, out_rows, rowStart);
    count_launch(2);
    rc = check_cuda(cudaGetLastError(), "csr kernels launch");
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");  // the scratch is freed on return
    return rc;
}

}  // extern "C"
