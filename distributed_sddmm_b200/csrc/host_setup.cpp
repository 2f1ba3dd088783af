// host_setup.cpp -- host-side (untimed) setup helpers exported through the C ABI:
// the seeded Erdos-Renyi tuple generator and the COO -> CSR conversion of one block.
//
// They replace, respectively, the CombBLAS Graph500 generator call in
// SpmatLocal::loadTuples (reference SpmatLocal.hpp:499-516) and the MKL inspector sequence
// in the CSRLocal constructor (reference SpmatLocal.hpp:78-188).  Neither CombBLAS nor MKL is
// used: both are plain C++ here.
#include "hnh_b200.h"
#include "launch.h"
#include "host_sort.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

// Rows [row_lo, row_hi) of the N x N (N = 2^logM) Erdos-Renyi matrix with seed `seed`:
// per row, nnz_per_row columns mix64(mix64(seed ^ row*ROWMUL) + k) mod N, sorted, unique.
// Output sorted by (row, col), values 1.0; independent of how rows are split over ranks.
// Returns the number of tuples written, or a negative HNH_E_* code (capacity too small ->
// HNH_E_INVALID; pass capacity >= (row_hi-row_lo)*nnz_per_row to be safe).
int64_t hnh_er_generate_host(int logM, int nnz_per_row, uint64_t seed, int64_t row_lo,
                             int64_t row_hi, uint64_t *rows_out, uint64_t *cols_out,
                             double *vals_out, int64_t capacity) {
    if (logM < 0 || logM > 40 || nnz_per_row < 0 || row_lo < 0 || row_hi < row_lo ||
        row_hi > ((int64_t)1 << logM))
        return hnh::set_error(HNH_E_INVALID, "hnh_er_generate_host: bad argument");
    const int64_t nrows = row_hi - row_lo;
    if (nrows == 0 || nnz_per_row == 0) return 0;
    if (!rows_out || !cols_out || !vals_out)
        return hnh::set_error(HNH_E_INVALID, "hnh_er_generate_host: null output");
    const uint64_t mask = ((uint64_t)1 << logM) - 1;
    std::vector<int64_t> counts((size_t)nrows + 1, 0);
    // pass 1: unique count per row
#pragma omp parallel
    {
        std::vector<uint64_t> tmp((size_t)nnz_per_row);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < nrows; i++) {
            const uint64_t base = mix64(seed ^ ((uint64_t)(row_lo + i) * 0xD1342543DE82EF95ull));
            for (int k = 0; k < nnz_per_row; k++) tmp[k] = mix64(base + (uint64_t)k) & mask;
            std::sort(tmp.begin(), tmp.end());
            counts[i + 1] = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
        }
    }
    for (int64_t i = 0; i < nrows; i++) counts[i + 1] += counts[i];
    const int64_t total = counts[nrows];
    if (total > capacity)
        return hnh::set_error(HNH_E_INVALID, "hnh_er_generate_host: capacity %lld < %lld",
                              (long long)capacity, (long long)total);
#pragma omp parallel
    {
        std::vector<uint64_t> tmp((size_t)nnz_per_row);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < nrows; i++) {
            const uint64_t base = mix64(seed ^ ((uint64_t)(row_lo + i) * 0xD1342543DE82EF95ull));
            for (int k = 0; k < nnz_per_row; k++) tmp[k] = mix64(base + (uint64_t)k) & mask;
            std::sort(tmp.begin(), tmp.end());
            const int64_t n = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
            int64_t o = counts[i];
            for (int64_t k = 0; k < n; k++, o++) {
                rows_out[o] = (uint64_t)(row_lo + i);
                cols_out[o] = tmp[k];
                vals_out[o] = 1.0;
            }
        }
    }
    return total;
}

// COO -> CSR of one block with optional transposition; the canonical order is stored-row
// ascending then input order (a stable counting sort) -- see DESIGN.md "CSR order".
// rowStart: out_rows+1 entries where out_rows = transpose ? cols : rows.  row_idx may be NULL.
int hnh_coo_to_csr_host(int64_t rows, int64_t cols, int64_t nnz, const uint64_t *r,
                        const uint64_t *c, const double *v, int transpose, int64_t *rowStart,
                        int64_t *col_idx, int64_t *row_idx, double *values) {
    if (rows < 0 || cols < 0 || nnz < 0 || !rowStart)
        return hnh::set_error(HNH_E_INVALID, "hnh_coo_to_csr_host: bad argument");
    if (nnz > 0 && (!r || !c || !v || !col_idx || !values))
        return hnh::set_error(HNH_E_INVALID, "hnh_coo_to_csr_host: null pointer");
    const int64_t out_rows = transpose ? cols : rows;
    const int64_t out_cols = transpose ? rows : cols;
    const uint64_t *sr = transpose ? c : r;  // stored row
    const uint64_t *sc = transpose ? r : c;  // stored col
    bool col_bad = false;
#pragma omp parallel for reduction(|| : col_bad)
    for (int64_t i = 0; i < nnz; i++) col_bad = col_bad || sc[i] >= (uint64_t)out_cols;
    const bool ok = !col_bad && hnh::stable_counting_sort(
        nnz, out_rows, [&](int64_t i) { return sr[i] < (uint64_t)out_rows ? (int64_t)sr[i] : (int64_t)-1; },
        [&](int64_t i, int64_t p) {
            col_idx[p] = (int64_t)sc[i];
            values[p] = v[i];
            if (row_idx) row_idx[p] = (int64_t)sr[i];
        },
        rowStart);
    if (!ok)
        return hnh::set_error(HNH_E_INVALID, "hnh_coo_to_csr_host: a coordinate lies outside the %lld x %lld block",
                              (long long)rows, (long long)cols);
    return HNH_OK;
}

}  // extern "C"
