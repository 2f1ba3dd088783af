// distributed_sparse.cpp -- non-template members of Distributed_Sparse (hnh/distributed_sparse.h).
#include "hnh/distributed_sparse.h"

#include <algorithm>
#include <iostream>

#include "hnh_b200.h"

using hnh::Runtime;

Distributed_Sparse::Distributed_Sparse(KernelImplementation *k) {
    auto world = hnh::Comm::world();
    proc_rank = world->rank();
    p = world->size();
    verbose = false;
    overlap = true;
    kernel = k;
    algorithm_name = "";
    M = N = R = -1;
    localArows = localAcols = localBrows = localBcols = -1;
    c = -1;
    r_split = false;
    superclass_constructor_sentinel = 3;
}

void Distributed_Sparse::check_initialized() {
    auto need = [](bool ok, const char *what) {
        if (!ok) throw hnh::Error(HNH_E_INVALID, string("Distributed_Sparse subclass left uninitialised: ") + what);
    };
    need(algorithm_name != "", "algorithm_name");
    need(!proc_grid_names.empty(), "proc_grid_names");
    need(!perf_counter_keys.empty(), "perf_counter_keys");
    need(M != -1 && N != -1 && R != -1, "M/N/R");
    need(localAcols != -1 && localBcols != -1 && localArows != -1 && localBrows != -1, "local dims");
    need(c >= 1, "c");
    need(superclass_constructor_sentinel == 3, "base constructor");
    need(!aSubmatrices.empty() && !bSubmatrices.empty(), "submatrix descriptors");
    need(S && ST && S->initialized && ST->initialized, "S/ST");
    need(S->coordinate_ownership_initialized && ST->coordinate_ownership_initialized, "coordinate ownership");
    need(!S->blockStarts.empty() && !ST->blockStarts.empty(), "blockStarts");
    need(S->csr_initialized && ST->csr_initialized, "CSR blocks");
}

json Distributed_Sparse::json_algorithm_info() {
    json j = json::object();
    j["alg_name"] = algorithm_name;
    j["m"] = (long long)M;
    j["n"] = (long long)N;
    j["nnz"] = (unsigned long long)S->dist_nnz;
    j["r"] = (long long)R;
    j["adjacency_mode"] = grid->adjacency;
    j["p"] = p;
    j["c"] = c;
    json interp = json::array(), vals = json::array();
    for (size_t i = 0; i < proc_grid_names.size(); i++) {
        interp.push_back(proc_grid_names[i]);
        vals.push_back(grid->dim_list[i]);
    }
    j["dim_interpretations"] = interp;
    j["dim_values"] = vals;
    // per-rank nonzero counts (load imbalance report)
    uint64_t mine[2] = {(uint64_t)(S->owned_coords_end - S->owned_coords_start),
                        (uint64_t)(ST->owned_coords_end - ST->owned_coords_start)};
    vector<uint64_t> all((size_t)2 * p);
    hnh::Comm::world()->host_allgather(mine, all.data(), sizeof(mine));
    json nnz = json::array(), nnz_t = json::array();
    for (int i = 0; i < p; i++) {
        nnz.push_back((unsigned long long)all[2 * i]);
        nnz_t.push_back((unsigned long long)all[2 * i + 1]);
    }
    j["nnz_procs"] = nnz;
    j["nnz_tpose_procs"] = nnz_t;
    // not in the reference's record: what actually carries the ring shifts of this object
    j["transport"] = hnh::Comm::world()->transport_name();
    j["peer_rings"] = (int)(peer_rings_.size() + sparse_rings_.size());
    j["ring"] = p == 1 ? "none" : (!peer_rings_.empty() || !sparse_rings_.empty())
                                      ? "peer_ring (copy engines into CUDA-IPC mapped peer slots, stream memory-op flags)"
                                      : "send/recv of the transport";
    return j;
}

void Distributed_Sparse::print_algorithm_info() {
    json j = json_algorithm_info();
    if (proc_rank == 0) cout << j.dump(4) << endl;
}

void Distributed_Sparse::reset_performance_timers() {
    for (auto &key : perf_counter_keys) {
        call_count[key] = 0;
        total_time[key] = 0.0;
    }
    timers_.reset();
}

static void require_key(const vector<string> &keys, const string &name) {
    if (find(keys.begin(), keys.end(), name) == keys.end())
        throw hnh::Error(HNH_E_INVALID, "Error, performance counter " + name + " not registered.");
}

void Distributed_Sparse::stop_clock_and_add(my_timer_t &start, string counter_name) {
    require_key(perf_counter_keys, counter_name);
    call_count[counter_name]++;
    total_time[counter_name] += stop_clock_get_elapsed(start);
}

void Distributed_Sparse::region_begin(const string &name, cudaStream_t s) {
    require_key(perf_counter_keys, name);
    timers_.start(name, s);
}

void Distributed_Sparse::region_end(const string &name, cudaStream_t s) {
    timers_.stop(name, s);
    call_count[name]++;
}

json Distributed_Sparse::json_perf_statistics() {
    json j = json::object();
    vector<double> vals;
    for (auto &key : perf_counter_keys) vals.push_back(total_time[key] + timers_.total_seconds(key));
    hnh::Comm::world()->host_allreduce_sum_f64(vals.data(), vals.size());
    for (size_t i = 0; i < perf_counter_keys.size(); i++) j[perf_counter_keys[i]] = vals[i] / p;
    return j;
}

void Distributed_Sparse::print_performance_statistics() {
    json info = json_algorithm_info();
    json stats = json_perf_statistics();
    if (proc_rank == 0) {
        cout << endl
             << "================================" << endl
             << "==== Performance Statistics ====" << endl
             << "================================" << endl
             << info.dump(4) << endl
             << stats.dump(4) << endl
             << "=================================" << endl;
    }
}

void Distributed_Sparse::fusedSpMM(DenseMatrix &localA, DenseMatrix &localB, VectorXd &Svalues, VectorXd &sddmm_buffer,
                                   MatMode mode) {
    if (mode == Amat) {
        algorithm(localA, localB, Svalues, &sddmm_buffer, k_sddmmA, true);
        localA.setZero();
        algorithm(localA, localB, sddmm_buffer, nullptr, k_spmmA, false);
    } else {
        algorithm(localA, localB, Svalues, &sddmm_buffer, k_sddmmB, true);
        localB.setZero();
        algorithm(localA, localB, sddmm_buffer, nullptr, k_spmmB, false);
    }
}

void Distributed_Sparse::dummyInitialize(DenseMatrix &loc, MatMode mode) {
    vector<DenseSubmatrix> &subs = (mode == Amat) ? aSubmatrices : bSubmatrices;
    vector<double> host((size_t)loc.size());
    size_t at = 0;
    for (const DenseSubmatrix &s : subs)
        for (int i = 0; i < s.rowCount; i++)
            for (int j = 0; j < s.colCount; j++) host[at++] = (double)(s.topRow + i) * (double)R + (double)(s.leftCol + j);
    if ((int64_t)at != loc.size()) throw hnh::Error(HNH_E_INVALID, "dummyInitialize: submatrices do not tile the local matrix");
    loc.copy_from_host(host.data());
}

void Distributed_Sparse::shiftDenseMatrix(BufferPair &buf, hnh::Comm &world, int send_dst, int /*tag*/, int recv_src) {
    Runtime &rt = Runtime::get();
    const int n = world.size();
    if (recv_src < 0) recv_src = pMod(2 * world.rank() - send_dst, n);
    const size_t bytes = sizeof(double) * (size_t)buf.getActive()->size();
    rt.chain(compute(), comm());
    world.sendrecv(buf.getActive()->data(), bytes, send_dst, buf.getPassive()->data(), bytes, recv_src, comm());
    rt.chain(comm(), compute());
    buf.swapActive();
}

double Distributed_Sparse::fingerprint(const DenseMatrix &m) {
    double v = m.squaredNorm();
    hnh::Comm::world()->host_allreduce_sum_f64(&v, 1);
    return v;
}

double Distributed_Sparse::fingerprint(const VectorXd &v) {
    double x = v.squaredNorm();
    hnh::Comm::world()->host_allreduce_sum_f64(&x, 1);
    return x;
}

void Distributed_Sparse::hadamard_values(VectorXd &result, VectorXd &SValues, SpmatLocal &m) {
    const int64_t total = (int64_t)m.blockStarts.back();
    if (SValues.size() != total)
        throw hnh::Error(HNH_E_INVALID, "SValues has " + to_string(SValues.size()) + " entries, the local sparse matrix " +
                                            to_string(total));
    if (result.size() != total) result.resize(total);
    for (size_t b = 0; b + 1 < m.blockStarts.size(); b++) {
        const int64_t off = (int64_t)m.blockStarts[b], n = (int64_t)m.blockStarts[b + 1] - off;
        if (n == 0) continue;
        if (m.csr_blocks[b])
            hnh::abi_check(hnh_hadamard_f64(result.data() + off, SValues.data() + off,
                                            m.csr_blocks[b]->getActive()->values.data(), n, compute()),
                           "hadamard");
        else
            hnh::cuda_check(cudaMemsetAsync(result.data() + off, 0, sizeof(double) * (size_t)n, compute()), "memset");
    }
}

hnh::PeerRing *Distributed_Sparse::peer_ring(std::shared_ptr<hnh::Comm> world, size_t bytes) {
    if (peer_ring_broken_ || !hnh::PeerRing::enabled() || world->size() < 2) return nullptr;
    auto key = std::make_pair(world.get(), bytes);
    auto it = peer_rings_.find(key);
    if (it != peer_rings_.end()) return it->second.get();
    try {
        peer_rings_[key].reset(new hnh::PeerRing(world, bytes));
    } catch (const hnh::Error &e) {
        // peer mapping unavailable (e.g. no NVLink/IPC): the NCCL ring below does the same job
        if (verbose) cout << "PeerRing unavailable, using NCCL send/recv: " << e.what() << endl;
        peer_rings_.erase(key);
        peer_ring_broken_ = true;
        return nullptr;
    }
    return peer_rings_[key].get();
}

hnh::PeerRing *Distributed_Sparse::sparse_ring(std::shared_ptr<hnh::Comm> world, CSRLocal *blk) {
    if (peer_ring_broken_ || !hnh::PeerRing::all_shifts() || world->size() < 2 || blk == nullptr) return nullptr;
    auto key = std::make_pair(world.get(), blk);
    auto it = sparse_rings_.find(key);
    if (it != sparse_rings_.end()) return it->second.get();
    try {
        sparse_rings_[key].reset(new hnh::PeerRing(world, blk->ring_buffers(), blk->active));
    } catch (const hnh::Error &e) {
        if (verbose) cout << "sparse PeerRing unavailable, using NCCL send/recv: " << e.what() << endl;
        sparse_rings_.erase(key);
        peer_ring_broken_ = true;
        return nullptr;
    }
    return sparse_rings_[key].get();
}

// logical buffers of CSRLocal::ring_buffers(): 0 = values, 1 = col_idx, 2 = rowStart
void Distributed_Sparse::sparse_push_early(hnh::PeerRing &pr, CSRLocal &blk, bool values_written) {
    const int k = 1 - blk.active;  // the downstream rank's passive handle (all ranks flip together)
    CSRHandle *h = blk.getActive();
    pr.begin_push(k, comm());
    const size_t nz = (size_t)blk.num_coords;
    if (nz) hnh::cuda_check(cudaMemcpyAsync(pr.dst_ptr(1, k), h->col_idx.data(), sizeof(int64_t) * nz, cudaMemcpyDeviceToDevice, comm()), "push col_idx");
    hnh::cuda_check(cudaMemcpyAsync(pr.dst_ptr(2, k), h->rowStart.data(), sizeof(int64_t) * (size_t)(blk.rows + 1), cudaMemcpyDeviceToDevice, comm()), "push rowStart");
    if (!values_written && nz)
        hnh::cuda_check(cudaMemcpyAsync(pr.dst_ptr(0, k), h->values.data(), sizeof(double) * nz, cudaMemcpyDeviceToDevice, comm()), "push values");
}

void Distributed_Sparse::sparse_push_late(hnh::PeerRing &pr, CSRLocal &blk, bool values_written, bool early_done, int64_t incoming) {
    const int a = blk.active, k = 1 - a;
    if (!early_done) sparse_push_early(pr, blk, false);  // everything in one go, after the kernel
    else if (values_written && blk.num_coords)
        hnh::cuda_check(cudaMemcpyAsync(pr.dst_ptr(0, k), blk.getActive()->values.data(), sizeof(double) * (size_t)blk.num_coords,
                                        cudaMemcpyDeviceToDevice, comm()), "push values");
    pr.end_push(k, comm());
    pr.release(a, comm());  // my handle `a` has been read by the kernel and by the push: upstream may refill it
    blk.shift_commit(incoming);
    pr.expect_arrival(k);
    pr.wait_arrival(k, comm());
}
