// host_sort.h -- stable parallel counting sort used by the (untimed) host setup path: COO -> CSR of a block
// (replacing the MKL inspector conversion, reference SpmatLocal.hpp:117-147) and the (column, row) ordering of
// redistributed tuples (replacing __gnu_parallel::sort with the column_major comparator, SpmatLocal.hpp:458).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <vector>

namespace hnh {

// Items 0..n-1 with key(i) in [0, nkeys).  Calls place(i, pos) with pos = the item's position in the stable order
// by key; returns false if a key was out of range (nothing is placed then).  starts (nkeys + 1 entries, optional)
// receives the first position of every key.
template <class KeyFn, class PlaceFn>
bool stable_counting_sort(int64_t n, int64_t nkeys, KeyFn key, PlaceFn place, int64_t *starts) {
    if (nkeys <= 0) return n == 0;
    int T = std::max(1, omp_get_max_threads());
    // bound the histogram memory (T * nkeys * 4 bytes) to ~256 MB and keep chunks reasonably large
    while (T > 1 && ((int64_t)T * nkeys * 4 > (256ll << 20) || n / T < (1 << 16))) T--;
    const int64_t chunk = (n + T - 1) / T;
    std::vector<uint32_t> counts((size_t)T * (size_t)nkeys, 0);
    bool bad = false;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(|| : bad)
    for (int t = 0; t < T; t++) {
        uint32_t *c = counts.data() + (size_t)t * (size_t)nkeys;
        const int64_t lo = std::min(n, chunk * t), hi = std::min(n, lo + chunk);
        for (int64_t i = lo; i < hi; i++) {
            const int64_t k = key(i);
            if (k < 0 || k >= nkeys) { bad = true; break; }
            c[k]++;
        }
    }
    if (bad) return false;
    // per-key totals -> exclusive prefix over keys -> per-(thread, key) start offsets
    std::vector<int64_t> base((size_t)nkeys + 1);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nkeys; k++) {
        int64_t s = 0;
        for (int t = 0; t < T; t++) s += counts[(size_t)t * (size_t)nkeys + (size_t)k];
        base[(size_t)k + 1] = s;
    }
    base[0] = 0;
    for (int64_t k = 0; k < nkeys; k++) base[(size_t)k + 1] += base[(size_t)k];
    if (starts) std::copy(base.begin(), base.end(), starts);
    std::vector<int64_t> offs((size_t)T * (size_t)nkeys);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nkeys; k++) {
        int64_t s = base[(size_t)k];
        for (int t = 0; t < T; t++) {
            offs[(size_t)t * (size_t)nkeys + (size_t)k] = s;
            s += counts[(size_t)t * (size_t)nkeys + (size_t)k];
        }
    }
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        int64_t *o = offs.data() + (size_t)t * (size_t)nkeys;
        const int64_t lo = std::min(n, chunk * t), hi = std::min(n, lo + chunk);
        for (int64_t i = lo; i < hi; i++) place(i, o[key(i)]++);
    }
    return true;
}

}  // namespace hnh
