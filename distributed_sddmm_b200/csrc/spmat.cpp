// spmat.cpp -- CSRLocal / SpmatLocal of hnh/SpmatLocal.hpp.
#include "hnh/SpmatLocal.hpp"

#include <parallel/algorithm>

#include <algorithm>
#include <fstream>
#include <numeric>
#include <sstream>

#include "hnh_b200.h"
#include "host_sort.h"

using hnh::abi_check;
using hnh::cuda_check;
using hnh::Runtime;

uint64_t SpmatLocal::er_seed = 0xC0FFEEull;

// ------------------------------------------------------------------ CSRLocal -------------
void CSRLocal::allocate(CSRHandle &h) {
    const size_t cap = (size_t)std::max<int64_t>(max_nnz, 1);
    h.values.resize(cap);
    h.col_idx.resize(cap);
    h.rowStart.resize((size_t)rows + 1);
    h.allocated = true;
    h.row_idx_valid = false;
}

CSRLocal::CSRLocal(int64_t rows_in, int64_t cols_in, int64_t max_nnz_in, spcoord_t *coords, int64_t n, bool tr)
    : rows(tr ? cols_in : rows_in), cols(tr ? rows_in : cols_in), max_nnz(max_nnz_in), num_coords(n),
      transpose(tr), active(0), buffer(new CSRHandle[2]) {
    if (n > max_nnz) throw hnh::Error(HNH_E_INVALID, "CSRLocal: num_coords > max_nnz");
    vector<int64_t> rowStart((size_t)rows + 1), col_idx((size_t)std::max<int64_t>(n, 1)), row_idx((size_t)std::max<int64_t>(n, 1));
    vector<double> values((size_t)std::max<int64_t>(n, 1));
    {
        // COO -> CSR straight from the tuples (stable counting sort by stored row)
        const uint64_t lim_r = (uint64_t)rows, lim_c = (uint64_t)cols;
        bool col_bad = false;
#pragma omp parallel for reduction(|| : col_bad)
        for (int64_t i = 0; i < n; i++) col_bad = col_bad || (tr ? coords[i].r : coords[i].c) >= lim_c;
        const bool ok = !col_bad && hnh::stable_counting_sort(
            n, rows,
            [&](int64_t i) {
                const uint64_t sr = tr ? coords[i].c : coords[i].r;
                return sr < lim_r ? (int64_t)sr : (int64_t)-1;
            },
            [&](int64_t i, int64_t p) {
                col_idx[p] = (int64_t)(tr ? coords[i].r : coords[i].c);
                row_idx[p] = (int64_t)(tr ? coords[i].c : coords[i].r);
                values[p] = coords[i].value;
            },
            rowStart.data());
        if (!ok) throw hnh::Error(HNH_E_INVALID, "CSRLocal: a coordinate lies outside the block");
    }
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        coords[i].r = (uint64_t)row_idx[i];
        coords[i].c = (uint64_t)col_idx[i];
        coords[i].value = values[i];
    }
    col_idx.resize((size_t)n);
    row_idx.resize((size_t)n);
    values.resize((size_t)n);
    host_.rowStart.swap(rowStart);
    host_.col_idx.swap(col_idx);
    host_.row_idx.swap(row_idx);
    host_.values.swap(values);
}

void CSRLocal::ensure_device() {
    if (on_device_) return;
    allocate(buffer[active]);
    cudaStream_t s = Runtime::get().compute_stream();
    buffer[active].rowStart.upload(host_.rowStart.data(), host_.rowStart.size(), s);
    if (num_coords > 0) {
        buffer[active].col_idx.upload(host_.col_idx.data(), (size_t)num_coords, s);
        buffer[active].values.upload(host_.values.data(), (size_t)num_coords, s);
    }
    cuda_check(cudaStreamSynchronize(s), "CSRLocal upload");
    host_ = HostCSR();
    on_device_ = true;
}

CSRLocal::~CSRLocal() { delete[] buffer; }

void CSRLocal::shiftCSR(int src, int dst, hnh::Comm &comm, int64_t nnz_to_receive, int /*tag*/, ShiftMode /*mode*/,
                        cudaStream_t s) {
    if (nnz_to_receive > std::max<int64_t>(max_nnz, 1))
        throw hnh::Error(HNH_E_INVALID, "shiftCSR: incoming block exceeds max_nnz");
    if (!on_device_) {  // setup-time shift of a block that has not been moved to HBM yet
        HostCSR in;
        in.rowStart.resize((size_t)rows + 1);
        in.col_idx.resize((size_t)nnz_to_receive);
        in.values.resize((size_t)nnz_to_receive);
        comm.host_sendrecv(host_.values.data(), sizeof(double) * (size_t)num_coords, dst, in.values.data(),
                           sizeof(double) * (size_t)nnz_to_receive, src);
        comm.host_sendrecv(host_.col_idx.data(), sizeof(int64_t) * (size_t)num_coords, dst, in.col_idx.data(),
                           sizeof(int64_t) * (size_t)nnz_to_receive, src);
        comm.host_sendrecv(host_.rowStart.data(), sizeof(int64_t) * (size_t)(rows + 1), dst, in.rowStart.data(),
                           sizeof(int64_t) * (size_t)(rows + 1), src);
        in.row_idx.resize((size_t)nnz_to_receive);
        for (int64_t r = 0; r < rows; r++)
            for (int64_t j = in.rowStart[r]; j < in.rowStart[r + 1]; j++) in.row_idx[j] = r;
        host_ = std::move(in);
        num_coords = nnz_to_receive;
        return;
    }
    if (!s) s = Runtime::get().comm_stream();
    shiftCSR_no_flip(src, dst, comm, nnz_to_receive, s);
    shift_commit(nnz_to_receive);
}

void CSRLocal::shiftCSR_no_flip(int src, int dst, hnh::Comm &comm, int64_t nnz_in, cudaStream_t s) {
    if (nnz_in > std::max<int64_t>(max_nnz, 1)) throw hnh::Error(HNH_E_INVALID, "shiftCSR: incoming block exceeds max_nnz");
    ensure_device();
    CSRHandle &snd = buffer[active];
    CSRHandle &rcv = buffer[1 - active];
    if (!rcv.allocated) allocate(rcv);
    hnh::Comm::Seg segs[3] = {
        {snd.values.data(), sizeof(double) * (size_t)num_coords, rcv.values.data(), sizeof(double) * (size_t)nnz_in},
        {snd.col_idx.data(), sizeof(int64_t) * (size_t)num_coords, rcv.col_idx.data(), sizeof(int64_t) * (size_t)nnz_in},
        {snd.rowStart.data(), sizeof(int64_t) * (size_t)(rows + 1), rcv.rowStart.data(), sizeof(int64_t) * (size_t)(rows + 1)},
    };
    comm.sendrecv_multi(segs, 3, dst, src, s);
}

void CSRLocal::shift_structure(int src, int dst, hnh::Comm &comm, int64_t nnz_in, cudaStream_t s) {
    if (nnz_in > std::max<int64_t>(max_nnz, 1)) throw hnh::Error(HNH_E_INVALID, "shiftCSR: incoming block exceeds max_nnz");
    ensure_device();
    CSRHandle &snd = buffer[active];
    CSRHandle &rcv = buffer[1 - active];
    if (!rcv.allocated) allocate(rcv);
    hnh::Comm::Seg segs[2] = {
        {snd.col_idx.data(), sizeof(int64_t) * (size_t)num_coords, rcv.col_idx.data(), sizeof(int64_t) * (size_t)nnz_in},
        {snd.rowStart.data(), sizeof(int64_t) * (size_t)(rows + 1), rcv.rowStart.data(), sizeof(int64_t) * (size_t)(rows + 1)},
    };
    comm.sendrecv_multi(segs, 2, dst, src, s);
}

void CSRLocal::shift_values(int src, int dst, hnh::Comm &comm, int64_t nnz_in, cudaStream_t s) {
    CSRHandle &snd = buffer[active];
    CSRHandle &rcv = buffer[1 - active];
    comm.sendrecv(snd.values.data(), sizeof(double) * (size_t)num_coords, dst, rcv.values.data(),
                  sizeof(double) * (size_t)nnz_in, src, s);
}

std::vector<std::array<void *, 2>> CSRLocal::ring_buffers() {
    ensure_device();
    for (int t = 0; t < 2; t++)
        if (!buffer[t].allocated) allocate(buffer[t]);
    return {{buffer[0].values.data(), buffer[1].values.data()},
            {buffer[0].col_idx.data(), buffer[1].col_idx.data()},
            {buffer[0].rowStart.data(), buffer[1].rowStart.data()}};
}

void CSRLocal::shift_commit(int64_t nnz_received) {
    buffer[1 - active].row_idx_valid = false;
    num_coords = nnz_received;
    active = 1 - active;
}

const int64_t *CSRLocal::row_idx_device() {
    CSRHandle *h = getActive();  // uploads if needed
    if (!h->row_idx_valid) {
        h->row_idx.resize((size_t)std::max<int64_t>(max_nnz, 1));
        abi_check(hnh_expand_row_idx(h->rowStart.data(), rows, num_coords, h->row_idx.data(),
                                     Runtime::get().compute_stream()),
                  "expand_row_idx");
        h->row_idx_valid = true;
    }
    return h->row_idx.data();
}

CSRLocal::HostCSR CSRLocal::to_host() {
    if (!on_device_) return host_;
    HostCSR out;
    cudaStream_t s = Runtime::get().compute_stream();
    Runtime::get().chain(Runtime::get().comm_stream(), s);
    const int64_t *ri = row_idx_device();
    CSRHandle *h = getActive();
    out.rowStart = h->rowStart.to_host((size_t)rows + 1, s);
    out.col_idx = h->col_idx.to_host((size_t)num_coords, s);
    out.values = h->values.to_host((size_t)num_coords, s);
    out.row_idx.resize((size_t)num_coords);
    if (num_coords) {
        cuda_check(cudaMemcpyAsync(out.row_idx.data(), ri, sizeof(int64_t) * (size_t)num_coords, cudaMemcpyDeviceToHost, s), "d2h");
        cuda_check(cudaStreamSynchronize(s), "sync");
    }
    return out;
}

// ------------------------------------------------------------------ SpmatLocal -----------
SpmatLocal::~SpmatLocal() {
    for (CSRLocal *b : csr_blocks) delete b;
}

void SpmatLocal::initializeCSRBlocks(int blockRows, int blockCols, int max_nnz, bool transpose) {
    if (max_nnz == -1) {
        for (size_t i = 0; i + 1 < blockStarts.size(); i++) {
            const int64_t n = (int64_t)(blockStarts[i + 1] - blockStarts[i]);
            csr_blocks.push_back(n > 0 ? new CSRLocal(blockRows, blockCols, n, coords.data() + blockStarts[i], n, transpose)
                                       : nullptr);
        }
    } else {
        const int64_t n = (int64_t)(blockStarts[1] - blockStarts[0]);
        csr_blocks.push_back(new CSRLocal(blockRows, blockCols, max_nnz, coords.data(), n, transpose));
    }
    csr_initialized = true;
}

void SpmatLocal::own_all_coordinates() {
    owned_coords_start = 0;
    owned_coords_end = (int)coords.size();
    layer_coords_start = {0, (int)coords.size()};
    layer_coords_sizes = {(int)coords.size()};
    coordinate_ownership_initialized = true;
}

void SpmatLocal::shard_across_layers(int num_layers, int current_layer) {
    layer_coords_start.clear();
    layer_coords_sizes.clear();
    divideIntoSegments((int)coords.size(), num_layers, layer_coords_start, layer_coords_sizes);
    owned_coords_start = layer_coords_start[current_layer];
    owned_coords_end = layer_coords_start[current_layer + 1];
    coordinate_ownership_initialized = true;
}

SpmatLocal *SpmatLocal::redistribute_nonzeros(NonzeroDistribution *dist, bool transpose, bool in_place) {
    hnh::Comm &comm = *dist->world;
    const int p = comm.size();
    const int64_t n = (int64_t)coords.size();

    // destination of every tuple, then a counting sort into per-destination segments
    vector<int> owner((size_t)n);
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) owner[i] = dist->getOwner((int64_t)coords[i].r, (int64_t)coords[i].c, transpose);
    // stable parallel counting sort of the tuples into per-destination segments
    vector<size_t> send_tuples((size_t)p, 0);
    vector<size_t> send_off((size_t)p + 1, 0);
    vector<spcoord_t> sendbuf((size_t)n);
    {
        vector<int64_t> starts((size_t)p + 1, 0);
        const bool ok = hnh::stable_counting_sort(
            n, p, [&](int64_t i) { return (int64_t)owner[i]; },
            [&](int64_t i, int64_t pos) {
                spcoord_t t;
                t.r = transpose ? coords[i].c : coords[i].r;
                t.c = transpose ? coords[i].r : coords[i].c;
                t.value = coords[i].value;
                sendbuf[(size_t)pos] = t;
            },
            starts.data());
        if (!ok) throw hnh::Error(HNH_E_INVALID, "redistribute_nonzeros: owner out of range");
        for (int i = 0; i <= p; i++) send_off[(size_t)i] = (size_t)starts[(size_t)i];
        for (int i = 0; i < p; i++) send_tuples[(size_t)i] = send_off[(size_t)i + 1] - send_off[(size_t)i];
    }
    vector<int>().swap(owner);

    // everyone learns the p x p count matrix (reference: MPI_Alltoall of counts, :425)
    vector<uint64_t> mine((size_t)p), all((size_t)p * p);
    for (int i = 0; i < p; i++) mine[i] = send_tuples[i];
    comm.host_allgather(mine.data(), all.data(), sizeof(uint64_t) * (size_t)p);
    const size_t B = sizeof(spcoord_t);
    vector<size_t> sb((size_t)p), sd((size_t)p), rb((size_t)p), rd((size_t)p);
    size_t total = 0;
    for (int i = 0; i < p; i++) {
        sb[i] = send_tuples[i] * B;
        sd[i] = send_off[i] * B;
        rb[i] = (size_t)all[(size_t)i * p + comm.rank()] * B;
        rd[i] = total;
        total += rb[i];
    }

    SpmatLocal *result = in_place ? this : new SpmatLocal();
    vector<spcoord_t> received(total / B);
    comm.host_alltoallv(sendbuf.data(), sb.data(), sd.data(), received.data(), rb.data(), rd.data());
    vector<spcoord_t>().swap(sendbuf);

    const uint64_t oldM = M, oldN = N;
    result->M = transpose ? oldN : oldM;
    result->N = transpose ? oldM : oldN;
    result->dist_nnz = dist_nnz;
    result->initialized = true;
    __gnu_parallel::sort(received.begin(), received.end(), column_major);
    result->coords.swap(received);
    return result;
}

void SpmatLocal::setTuples(uint64_t M_, uint64_t N_, const uint64_t *r, const uint64_t *c, const double *v, int64_t n) {
    coords.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) coords[i] = spcoord_t{r[i], c[i], v[i]};
    M = M_;
    N = N_;
    double cnt = (double)n;
    hnh::Comm::world()->host_allreduce_sum_f64(&cnt, 1);
    dist_nnz = (uint64_t)cnt;
    initialized = true;
}

static void read_matrix_market(const string &filename, int rank, int p, vector<spcoord_t> &coords, uint64_t &M,
                               uint64_t &N) {
    std::ifstream in(filename);
    if (!in) throw hnh::Error(HNH_E_INVALID, "loadTuples: cannot open " + filename);
    string line;
    bool pattern = false, symmetric = false;
    if (!std::getline(in, line) || line.rfind("%%MatrixMarket", 0) != 0)
        throw hnh::Error(HNH_E_INVALID, "loadTuples: not a MatrixMarket file");
    pattern = line.find("pattern") != string::npos;
    symmetric = line.find("symmetric") != string::npos;
    while (std::getline(in, line) && !line.empty() && line[0] == '%') {}
    uint64_t nnz = 0;
    {
        std::istringstream hs(line);
        hs >> M >> N >> nnz;
    }
    uint64_t idx = 0;
    for (uint64_t e = 0; e < nnz && std::getline(in, line); e++) {
        std::istringstream ls(line);
        uint64_t r, c;
        double v = 1.0;
        ls >> r >> c;
        if (!pattern) ls >> v;
        if ((idx++ % (uint64_t)p) == (uint64_t)rank) coords.push_back(spcoord_t{r - 1, c - 1, v});
        if (symmetric && r != c && (idx++ % (uint64_t)p) == (uint64_t)rank) coords.push_back(spcoord_t{c - 1, r - 1, v});
    }
}

void SpmatLocal::loadTuples(bool readFromFile, int logM, int nnz_per_row, string filename) {
    auto world = hnh::Comm::world();
    const int p = world->size(), rank = world->rank();
    coords.clear();
    if (readFromFile) {
        read_matrix_market(filename, rank, p, coords, M, N);
    } else {
        M = N = (uint64_t)1 << logM;
        const int64_t per = (int64_t)(M / (uint64_t)p);
        const int64_t lo = per * rank, hi = (rank == p - 1) ? (int64_t)M : per * (rank + 1);
        const int64_t cap = (hi - lo) * nnz_per_row;
        vector<uint64_t> r((size_t)std::max<int64_t>(cap, 1)), c((size_t)std::max<int64_t>(cap, 1));
        vector<double> v((size_t)std::max<int64_t>(cap, 1));
        const int64_t n = hnh_er_generate_host(logM, nnz_per_row, er_seed, lo, hi, r.data(), c.data(), v.data(), cap);
        if (n < 0) abi_check((int)n, "ER generator");
        coords.resize((size_t)n);
#pragma omp parallel for
        for (int64_t i = 0; i < n; i++) coords[i] = spcoord_t{r[i], c[i], v[i]};
    }
    double cnt = (double)coords.size();
    world->host_allreduce_sum_f64(&cnt, 1);
    dist_nnz = (uint64_t)cnt;
    initialized = true;
}

void SpmatLocal::divideIntoBlockCols(int blockWidth, int targetDivisions, bool modIndex) {
    // coords are sorted column-major: block k starts at the first tuple with c >= k*blockWidth
    blockStarts.assign((size_t)targetDivisions + 1, coords.size());
    for (int k = 0; k <= targetDivisions; k++) {
        spcoord_t key{0, (uint64_t)k * (uint64_t)blockWidth, 0.0};
        auto it = std::lower_bound(coords.begin(), coords.end(), key,
                                   [](const spcoord_t &a, const spcoord_t &b) { return a.c < b.c; });
        blockStarts[k] = (uint64_t)(it - coords.begin());
    }
    if (blockStarts[targetDivisions] != coords.size())
        throw hnh::Error(HNH_E_INVALID, "divideIntoBlockCols: column index beyond targetDivisions*blockWidth");
    if (modIndex) {
#pragma omp parallel for
        for (int64_t i = 0; i < (int64_t)coords.size(); i++) coords[i].c %= (uint64_t)blockWidth;
    }
}

void SpmatLocal::monolithBlockColumn() {
    blockStarts.clear();
    blockStarts.push_back(0);
    blockStarts.push_back(coords.size());
}

void SpmatLocal::setCSRValues(VectorXd &values) {
    cudaStream_t s = Runtime::get().compute_stream();
    for (size_t i = 0; i + 1 < blockStarts.size(); i++) {
        if (!csr_blocks[i]) continue;
        const size_t n = blockStarts[i + 1] - blockStarts[i];
        if (n)
            cuda_check(cudaMemcpyAsync(csr_blocks[i]->getActive()->values.data(), values.data() + blockStarts[i],
                                       sizeof(double) * n, cudaMemcpyDeviceToDevice, s),
                       "setCSRValues");
    }
}

void SpmatLocal::getCSRValues(VectorXd &out) {
    cudaStream_t s = Runtime::get().compute_stream();
    const int64_t total = (int64_t)blockStarts.back();
    if (out.size() != total) out.resize(total);
    for (size_t i = 0; i + 1 < blockStarts.size(); i++) {
        const size_t n = blockStarts[i + 1] - blockStarts[i];
        if (!n) continue;
        if (csr_blocks[i])
            cuda_check(cudaMemcpyAsync(out.data() + blockStarts[i], csr_blocks[i]->getActive()->values.data(),
                                       sizeof(double) * n, cudaMemcpyDeviceToDevice, s),
                       "getCSRValues");
        else
            cuda_check(cudaMemsetAsync(out.data() + blockStarts[i], 0, sizeof(double) * n, s), "getCSRValues");
    }
}

VectorXd SpmatLocal::getCSRValues() {
    VectorXd v;
    getCSRValues(v);
    return v;
}

void SpmatLocal::setValuesConstant(double cval) {
    cudaStream_t s = Runtime::get().compute_stream();
    for (size_t i = 0; i + 1 < blockStarts.size(); i++) {
        if (!csr_blocks[i]) continue;
        const int64_t n = (int64_t)(blockStarts[i + 1] - blockStarts[i]);
        abi_check(hnh_fill_f64(csr_blocks[i]->getActive()->values.data(), n, cval, s), "setValuesConstant");
    }
}
