// spmat.cpp -- CSRLocal / SpmatLocal of hnh/SpmatLocal.hpp.
#include "hnh/SpmatLocal.hpp"

#include <parallel/algorithm>

#include <algorithm>
#include <fstream>
#include <memory>
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

CSRLocal::CSRLocal(int64_t rows_in, int64_t cols_in, int64_t max_nnz_in, const uint64_t *d_r, const uint64_t *d_c, const double *d_v,
                   int64_t n, bool tr)
    : rows(tr ? cols_in : rows_in), cols(tr ? rows_in : cols_in), max_nnz(max_nnz_in), num_coords(n), transpose(tr), active(0),
      buffer(new CSRHandle[2]) {
    if (n > max_nnz) {
        delete[] buffer;
        throw hnh::Error(HNH_E_INVALID, "CSRLocal: num_coords > max_nnz");
    }
    try {
        allocate(buffer[active]);
        CSRHandle &h = buffer[active];
        abi_check(hnh_coo_to_csr_device(rows_in, cols_in, n, d_r, d_c, d_v, tr ? 1 : 0, h.rowStart.data(), h.col_idx.data(), nullptr,
                                        h.values.data(), Runtime::get().compute_stream()),
                  "hnh_coo_to_csr_device");
    } catch (...) {
        delete[] buffer;
        throw;
    }
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
    hnh::SetupPhase ph(tuples_on_device_ ? "COO -> CSR blocks (device)" : "COO -> CSR blocks (host)");
    auto make = [&](int64_t off, int64_t n, int64_t cap) -> CSRLocal * {
        if (tuples_on_device_)
            return new CSRLocal(blockRows, blockCols, cap, d_r_.data() + off, d_c_.data() + off, d_v_.data() + off, n, transpose);
        return new CSRLocal(blockRows, blockCols, cap, coords.data() + off, n, transpose);
    };
    if (max_nnz == -1) {
        for (size_t i = 0; i + 1 < blockStarts.size(); i++) {
            const int64_t n = (int64_t)(blockStarts[i + 1] - blockStarts[i]);
            csr_blocks.push_back(n > 0 ? make((int64_t)blockStarts[i], n, n) : nullptr);
        }
    } else {
        const int64_t n = (int64_t)(blockStarts[1] - blockStarts[0]);
        csr_blocks.push_back(make(0, n, max_nnz));
    }
    csr_initialized = true;
}

void SpmatLocal::own_all_coordinates() {
    const int n = (int)local_tuple_count();
    owned_coords_start = 0;
    owned_coords_end = n;
    layer_coords_start = {0, n};
    layer_coords_sizes = {n};
    coordinate_ownership_initialized = true;
}

void SpmatLocal::shard_across_layers(int num_layers, int current_layer) {
    layer_coords_start.clear();
    layer_coords_sizes.clear();
    divideIntoSegments((int)local_tuple_count(), num_layers, layer_coords_start, layer_coords_sizes);
    owned_coords_start = layer_coords_start[current_layer];
    owned_coords_end = layer_coords_start[current_layer + 1];
    coordinate_ownership_initialized = true;
}

SpmatLocal *SpmatLocal::redistribute_nonzeros(NonzeroDistribution *dist, bool transpose, bool in_place) {
    hnh::SetupPhase whole("redistribute (total)");
    if (use_device_setup()) return redistribute_nonzeros_device(dist, transpose, in_place);
    tuples_to_host();
    hnh::Comm &comm = *dist->world;
    const int p = comm.size();
    const int64_t n = (int64_t)coords.size();

    // destination of every tuple, then a counting sort into per-destination segments
    std::unique_ptr<hnh::SetupPhase> ph(new hnh::SetupPhase("redistribute: owner + bucket (host)"));
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

    ph.reset(new hnh::SetupPhase("redistribute: exchange"));
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
    ph.reset(new hnh::SetupPhase("redistribute: sort by (col, row) (host)"));
    __gnu_parallel::sort(received.begin(), received.end(), column_major);
    ph.reset();
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
        dist_nnz = M * (uint64_t)nnz_per_row;  // upper bound, for the (collective) choice of the setup path
        if (use_device_setup() && nnz_per_row <= 128) {
            hnh::SetupPhase ph("generate tuples (device)");
            d_r_.resize((size_t)std::max<int64_t>(cap, 1));
            d_c_.resize((size_t)std::max<int64_t>(cap, 1));
            d_v_.resize((size_t)std::max<int64_t>(cap, 1));
            const int64_t n = hnh_er_generate_device(logM, nnz_per_row, er_seed, lo, hi, d_r_.data(), d_c_.data(), d_v_.data(), cap,
                                                     Runtime::get().compute_stream());
            if (n < 0) abi_check((int)n, "ER generator (device)");
            d_n_ = n;
            tuples_on_device_ = true;
            double cnt = (double)n;
            world->host_allreduce_sum_f64(&cnt, 1);
            dist_nnz = (uint64_t)cnt;
            initialized = true;
            return;
        }
        hnh::SetupPhase ph("generate tuples (host)");
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
    if (tuples_on_device_) {
        vector<int64_t> starts((size_t)targetDivisions + 1);
        abi_check(hnh_tuples_block_starts_device(d_c_.data(), d_n_, (uint64_t)blockWidth, targetDivisions, starts.data(),
                                                 Runtime::get().compute_stream()), "hnh_tuples_block_starts_device");
        blockStarts.assign(starts.begin(), starts.end());
        if ((int64_t)blockStarts[(size_t)targetDivisions] != d_n_)
            throw hnh::Error(HNH_E_INVALID, "divideIntoBlockCols: column index beyond targetDivisions*blockWidth");
        if (modIndex) mod_coordinates(0, (uint64_t)blockWidth);
        return;
    }
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
    blockStarts.push_back((uint64_t)local_tuple_count());
}

// ------------------------------------------------------------------ device-resident tuples ---
bool SpmatLocal::use_device_setup() const {
    const int p = hnh::Comm::world_initialised() ? hnh::Comm::world()->size() : 1;
    // per-call limit of the CUB-based helpers: fewer than 2^31 tuples on one rank (dist_nnz bounds every rank's share)
    return dist_nnz < ((uint64_t)1 << 31) && hnh::device_setup_enabled((int64_t)(dist_nnz / (uint64_t)p));
}

// SoA copy of host tuples in HBM (the host vector is left alone)
static void upload_tuples(const vector<spcoord_t> &coords, hnh::DeviceBuffer<uint64_t> &d_r, hnh::DeviceBuffer<uint64_t> &d_c,
                          hnh::DeviceBuffer<double> &d_v) {
    const int64_t n = (int64_t)coords.size();
    vector<uint64_t> r((size_t)std::max<int64_t>(n, 1)), c((size_t)std::max<int64_t>(n, 1));
    vector<double> v((size_t)std::max<int64_t>(n, 1));
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        r[i] = coords[i].r;
        c[i] = coords[i].c;
        v[i] = coords[i].value;
    }
    cudaStream_t s = Runtime::get().compute_stream();
    d_r.resize((size_t)std::max<int64_t>(n, 1));
    d_c.resize((size_t)std::max<int64_t>(n, 1));
    d_v.resize((size_t)std::max<int64_t>(n, 1));
    d_r.upload(r.data(), (size_t)n, s);
    d_c.upload(c.data(), (size_t)n, s);
    d_v.upload(v.data(), (size_t)n, s);
    cuda_check(cudaStreamSynchronize(s), "upload_tuples");
}

void SpmatLocal::tuples_to_device() {
    if (tuples_on_device_) return;
    hnh::SetupPhase ph("tuples host -> device");
    upload_tuples(coords, d_r_, d_c_, d_v_);
    d_n_ = (int64_t)coords.size();
    tuples_on_device_ = true;
    vector<spcoord_t>().swap(coords);
}

void SpmatLocal::tuples_to_host() {
    if (!tuples_on_device_) return;
    hnh::SetupPhase ph("tuples device -> host");
    cudaStream_t s = Runtime::get().compute_stream();
    vector<uint64_t> r = d_r_.to_host((size_t)d_n_, s), c = d_c_.to_host((size_t)d_n_, s);
    vector<double> v = d_v_.to_host((size_t)d_n_, s);
    coords.resize((size_t)d_n_);
#pragma omp parallel for
    for (int64_t i = 0; i < d_n_; i++) coords[i] = spcoord_t{r[i], c[i], v[i]};
    d_r_.release();
    d_c_.release();
    d_v_.release();
    d_n_ = 0;
    tuples_on_device_ = false;
}

void SpmatLocal::mod_coordinates(uint64_t mod_r, uint64_t mod_c) {
    if (tuples_on_device_) {
        abi_check(hnh_tuples_mod_device(d_r_.data(), d_c_.data(), d_n_, mod_r, mod_c, Runtime::get().compute_stream()),
                  "hnh_tuples_mod_device");
        return;
    }
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)coords.size(); i++) {
        if (mod_r) coords[i].r %= mod_r;
        if (mod_c) coords[i].c %= mod_c;
    }
}

void SpmatLocal::release_tuples() {
    vector<spcoord_t>().swap(coords);
    if (tuples_on_device_) {
        // the CSR blocks were built from these arrays on the compute stream; the caching allocator reuses on that stream
        d_r_.release();
        d_c_.release();
        d_v_.release();
        d_n_ = 0;
        tuples_on_device_ = false;
    }
}

// redistribute_nonzeros with the tuples in HBM from start to finish: owner lookup + stable bucketing (CUB radix sort),
// device all-to-all (NCCL send/recv groups), (col, row) sort (two stable radix passes).  Same result as the host path:
// per source rank the tuples keep their order, the received set is ordered by (col, row).
SpmatLocal *SpmatLocal::redistribute_nonzeros_device(NonzeroDistribution *dist, bool transpose, bool in_place) {
    hnh::Comm &comm = *dist->world;
    const int p = comm.size();
    cudaStream_t s = Runtime::get().compute_stream();
    std::unique_ptr<hnh::SetupPhase> ph(new hnh::SetupPhase("redistribute: owner + bucket (device)"));

    // source tuples: this object's device copy, or a temporary upload of its host tuples (the source stays as it is)
    hnh::DeviceBuffer<uint64_t> tmp_r, tmp_c;
    hnh::DeviceBuffer<double> tmp_v;
    if (!tuples_on_device_) upload_tuples(coords, tmp_r, tmp_c, tmp_v);
    const uint64_t *src_r = tuples_on_device_ ? d_r_.data() : tmp_r.data(), *src_c = tuples_on_device_ ? d_c_.data() : tmp_c.data();
    const double *src_v = tuples_on_device_ ? d_v_.data() : tmp_v.data();
    const int64_t n = local_tuple_count();
    const uint64_t Mr = transpose ? N : M, Nr = transpose ? M : N;  // shape the distribution indexes into / of the result
    const int64_t row_blocks = std::max<int64_t>(1, (int64_t)((Mr + (uint64_t)dist->rows_in_block - 1) / (uint64_t)dist->rows_in_block));
    const int64_t col_blocks = std::max<int64_t>(1, (int64_t)((Nr + (uint64_t)dist->cols_in_block - 1) / (uint64_t)dist->cols_in_block));
    const vector<int> table = dist->owner_table(row_blocks, col_blocks);
    for (int o : table)
        if (o < 0 || o >= p) throw hnh::Error(HNH_E_INVALID, "redistribute_nonzeros: owner out of range");

    hnh::DeviceBuffer<uint64_t> sr((size_t)std::max<int64_t>(n, 1)), sc((size_t)std::max<int64_t>(n, 1));
    hnh::DeviceBuffer<double> sv((size_t)std::max<int64_t>(n, 1));
    vector<int64_t> starts((size_t)p + 1, 0);
    abi_check(hnh_tuples_bucket_by_owner_device(src_r, src_c, src_v, n, transpose ? 1 : 0,
                                                dist->rows_in_block, dist->cols_in_block, table.data(), row_blocks, col_blocks, p,
                                                sr.data(), sc.data(), sv.data(), starts.data(), s),
              "hnh_tuples_bucket_by_owner_device");
    tmp_r.release();
    tmp_c.release();
    tmp_v.release();

    ph.reset(new hnh::SetupPhase("redistribute: exchange"));
    vector<uint64_t> mine((size_t)p), all((size_t)p * p);
    for (int i = 0; i < p; i++) mine[i] = (uint64_t)(starts[(size_t)i + 1] - starts[(size_t)i]);
    comm.host_allgather(mine.data(), all.data(), sizeof(uint64_t) * (size_t)p);
    vector<size_t> sb((size_t)p), sd((size_t)p), rb((size_t)p), rd((size_t)p);
    size_t total = 0;
    for (int i = 0; i < p; i++) {
        sb[i] = (size_t)mine[i] * 8;
        sd[i] = (size_t)starts[i] * 8;
        rb[i] = (size_t)all[(size_t)i * p + comm.rank()] * 8;
        rd[i] = total;
        total += rb[i];
    }
    const int64_t nr = (int64_t)(total / 8);
    SpmatLocal *result = in_place ? this : new SpmatLocal();
    hnh::DeviceBuffer<uint64_t> rr((size_t)std::max<int64_t>(nr, 1)), rc((size_t)std::max<int64_t>(nr, 1));
    hnh::DeviceBuffer<double> rv((size_t)std::max<int64_t>(nr, 1));
    comm.alltoallv(sr.data(), sb.data(), sd.data(), rr.data(), rb.data(), rd.data(), s);
    comm.alltoallv(sc.data(), sb.data(), sd.data(), rc.data(), rb.data(), rd.data(), s);
    comm.alltoallv(sv.data(), sb.data(), sd.data(), rv.data(), rb.data(), rd.data(), s);
    cuda_check(cudaStreamSynchronize(s), "redistribute exchange");

    ph.reset(new hnh::SetupPhase("redistribute: sort by (col, row) (device)"));
    abi_check(hnh_tuples_sort_colmajor_device(rr.data(), rc.data(), rv.data(), nr, std::max(M, N), std::max(M, N), s),
              "hnh_tuples_sort_colmajor_device");
    ph.reset();

    const uint64_t oldM = M, oldN = N, nnz = dist_nnz;
    vector<spcoord_t>().swap(result->coords);
    result->d_r_ = std::move(rr);
    result->d_c_ = std::move(rc);
    result->d_v_ = std::move(rv);
    result->d_n_ = nr;
    result->tuples_on_device_ = true;
    result->M = transpose ? oldN : oldM;
    result->N = transpose ? oldM : oldN;
    result->dist_nnz = nnz;
    result->initialized = true;
    return result;
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
