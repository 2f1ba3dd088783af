// oracle/shims/CombBLAS/CombBLAS.h -- TEST INFRASTRUCTURE.  Just enough of the CombBLAS names that
// the reference's common.h / SpmatLocal.hpp mention (SpmatLocal.hpp:358-370,467-533) for those
// files to compile UNMODIFIED in oracle/_ref.  CombBLAS is not in this image.
//
// The reference uses CombBLAS for two things only: reading a MatrixMarket file and generating
// the R-MAT / Erdos-Renyi edge list (Graph500 generator with a uniform initiator).  Here
// DistEdgeList::GenGraph500Data produces this repo's seeded Erdos-Renyi tuples instead
// (the same generator as hnh_er_generate_host / oracle.er_tuples, seed = hnh_shim_er_seed), each
// rank its 1-D row slice with LOCAL row indices -- loadTuples then adds rank * (M / p) exactly as
// it does for CombBLAS output (SpmatLocal.hpp:526-529).
#pragma once
#include <mpi.h>

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

extern uint64_t hnh_shim_er_seed;

namespace combblas {

class CommGrid {
public:
    MPI_Comm world;
    int nprocs, rank;
    CommGrid(MPI_Comm w, int rows, int /*cols*/) : world(w), nprocs(rows) { MPI_Comm_rank(w, &rank); }
};

template <typename T>
struct maximum {
    T operator()(const T &a, const T &b) const { return a > b ? a : b; }
};

template <typename IT, typename NT>
class SpDCCols {};

template <typename IT>
class DistEdgeList {
public:
    std::shared_ptr<CommGrid> grid;
    std::vector<std::tuple<int64_t, int64_t, int>> edges;  // local rows
    int64_t global_rows = 0;
    explicit DistEdgeList(std::shared_ptr<CommGrid> g) : grid(g) {}
    static uint64_t mix64(uint64_t z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void GenGraph500Data(double * /*initiator*/, unsigned long scale, int edgefactor) {
        const int64_t N = (int64_t)1 << scale;
        global_rows = N;
        const int p = grid->nprocs, r = grid->rank;
        const int64_t per = N / p, lo = per * r, hi = (r == p - 1) ? N : per * (r + 1);
        std::vector<uint64_t> tmp((size_t)edgefactor);
        for (int64_t i = lo; i < hi; i++) {
            const uint64_t base = mix64(hnh_shim_er_seed ^ ((uint64_t)i * 0xD1342543DE82EF95ull));
            for (int k = 0; k < edgefactor; k++) tmp[(size_t)k] = mix64(base + (uint64_t)k) & (uint64_t)(N - 1);
            std::sort(tmp.begin(), tmp.end());
            const size_t n = (size_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
            for (size_t k = 0; k < n; k++) edges.emplace_back(i - lo, (int64_t)tmp[k], 1);
        }
    }
};
template <typename IT>
void PermEdges(DistEdgeList<IT> &) {}
template <typename IT>
void RenameVertices(DistEdgeList<IT> &) {}

template <typename IT, typename NT>
class SpTuples {
public:
    std::tuple<IT, IT, NT> *tuples = nullptr;
    int64_t nnz = 0;
    std::vector<std::tuple<IT, IT, NT>> store;
    SpTuples() {}
    SpTuples(const SpTuples &o) : nnz(o.nnz), store(o.store) { tuples = store.data(); }
    int64_t getnnz() const { return nnz; }
};

template <typename IT, typename NT, typename DER>
class SpParMat {
public:
    std::shared_ptr<CommGrid> grid;
    std::vector<std::tuple<int64_t, int64_t, int>> local;  // local rows, global cols
    int64_t nrow = 0, ncol = 0, total = 0;
    explicit SpParMat(std::shared_ptr<CommGrid> g) : grid(g) {}
    SpParMat(DistEdgeList<int64_t> &del, bool /*removeloops*/) : grid(del.grid), local(del.edges) {
        nrow = ncol = del.global_rows;
        long mine = (long)local.size(), all = 0;
        MPI_Allreduce(&mine, &all, 1, MPI_LONG, MPI_SUM, grid->world);
        total = all;
    }
    // coordinate MatrixMarket; rows are dealt to ranks in contiguous slices like a p x 1 grid
    template <typename OP>
    void ParallelReadMM(const std::string &filename, bool /*onebased*/, OP) {
        std::ifstream in(filename);
        std::string line;
        std::getline(in, line);
        const bool pattern = line.find("pattern") != std::string::npos;
        while (std::getline(in, line) && !line.empty() && line[0] == '%') {}
        int64_t nz = 0;
        { std::istringstream hs(line); hs >> nrow >> ncol >> nz; }
        const int p = grid->nprocs, r = grid->rank;
        const int64_t per = nrow / p, lo = per * r, hi = (r == p - 1) ? nrow : per * (r + 1);
        for (int64_t e = 0; e < nz && std::getline(in, line); e++) {
            std::istringstream ls(line);
            int64_t i, j; double v = 1.0;
            ls >> i >> j;
            if (!pattern) ls >> v;
            if (i - 1 >= lo && i - 1 < hi) local.emplace_back(i - 1 - lo, j - 1, (int)v);
        }
        total = nz;
    }
    int64_t getnnz() const { return total; }
    int64_t getnrow() const { return nrow; }
    int64_t getncol() const { return ncol; }
    SpTuples<int64_t, int> seq() const {
        SpTuples<int64_t, int> t;
        t.store = local;
        t.tuples = t.store.data();
        t.nnz = (int64_t)t.store.size();
        return t;
    }
};

}  // namespace combblas
