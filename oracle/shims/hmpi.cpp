// oracle/shims/hmpi.cpp -- TEST INFRASTRUCTURE.  Threads-as-ranks implementation of oracle/shims/mpi.h.
#include "mpi.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace {

struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    long gen = 0;
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        long g = gen;
        if (++waiting == n) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

struct Message {
    int src, tag;
    long epoch;  // -1, or the sender's MPI_Sendrecv call number on this (comm, tag)
    std::vector<char> data;
};

}  // namespace

struct hmpi_request {
    bool is_recv = false;
    void *buf = nullptr;
    size_t bytes = 0;
    int src = 0, tag = 0;
    long epoch = -1;
    MPI_Comm comm = nullptr;
    bool complete = false;  // guarded by comm->mail_mu
    MPI_Status status{0, 0, 0};
};

struct hmpi_comm {
    int size = 0;
    std::vector<int> world_ranks;  // member -> world rank
    Barrier bar;
    std::vector<const void *> slot;  // one pointer per member, for collectives
    std::vector<const void *> slot2;
    // mailboxes: per destination member
    std::mutex mail_mu;
    std::condition_variable mail_cv;
    std::vector<std::deque<Message>> mail;
    // receives posted with MPI_Irecv that have not been matched yet, per destination, in
    // posting order: a later send completes them directly (like a real MPI progress engine --
    // the reference posts Irecvs it never waits on, SpmatLocal.hpp:248-255)
    std::vector<std::deque<hmpi_request *>> pending;
    // MPI_Sendrecv call counters per member and tag.  The reference receives its ring shifts
    // from MPI_ANY_SOURCE with one tag for every phase (distributed_sparse.h:351-361), which is
    // only unambiguous if no rank runs a whole phase ahead -- true on a real cluster by timing,
    // not with eager thread-ranks.  Pairing the k-th Sendrecv of every member of a communicator
    // (SPMD: all members make the same sequence of calls) realises the intended matching.
    std::vector<std::map<int, long>> sendrecv_calls;
    explicit hmpi_comm(int n)
        : size(n), slot((size_t)n), slot2((size_t)n), mail((size_t)n), pending((size_t)n), sendrecv_calls((size_t)n) {
        bar.n = n;
    }
};

namespace {

hmpi_comm *g_world = nullptr;
thread_local int tl_world_rank = -1;
std::mutex g_type_mu;
std::vector<size_t> g_struct_sizes;  // id 100 + i

size_t type_size(MPI_Datatype t) {
    switch (t) {
        case MPI_CHAR: return 1;
        case MPI_INT: return 4;
        case MPI_FLOAT: return 4;
        case MPI_LONG: case MPI_DOUBLE: case MPI_UINT64_T: case MPI_UNSIGNED_LONG: return 8;
        default: {
            std::lock_guard<std::mutex> lk(g_type_mu);
            if (t >= 100 && (size_t)(t - 100) < g_struct_sizes.size()) return g_struct_sizes[(size_t)(t - 100)];
            fprintf(stderr, "hmpi: unknown datatype %d\n", t);
            abort();
        }
    }
}

int my_rank(hmpi_comm *c) {
    for (int i = 0; i < c->size; i++)
        if (c->world_ranks[(size_t)i] == tl_world_rank) return i;
    fprintf(stderr, "hmpi: calling thread (world rank %d) is not a member of this communicator\n", tl_world_rank);
    abort();
}

// publish a pointer, let everybody see everybody's; caller must call done(c) after reading
void publish(hmpi_comm *c, int me, const void *p, const void *p2 = nullptr) {
    c->slot[(size_t)me] = p;
    c->slot2[(size_t)me] = p2;
    c->bar.wait();
}
void done(hmpi_comm *c) { c->bar.wait(); }

template <typename T>
void sum_into(T *dst, const T *src, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] += src[i];
}
void reduce_sum(MPI_Datatype t, void *dst, const void *src, size_t n) {
    switch (t) {
        case MPI_DOUBLE: sum_into((double *)dst, (const double *)src, n); break;
        case MPI_INT: sum_into((int *)dst, (const int *)src, n); break;
        case MPI_LONG: sum_into((long *)dst, (const long *)src, n); break;
        case MPI_UINT64_T: case MPI_UNSIGNED_LONG: sum_into((uint64_t *)dst, (const uint64_t *)src, n); break;
        case MPI_FLOAT: sum_into((float *)dst, (const float *)src, n); break;
        default: fprintf(stderr, "hmpi: reduction on datatype %d\n", t); abort();
    }
}

void post(hmpi_comm *c, int me, int dst, int tag, const void *buf, size_t bytes, long epoch = -1) {
    {
        std::lock_guard<std::mutex> lk(c->mail_mu);
        auto &pend = c->pending[(size_t)dst];
        for (auto it = pend.begin(); it != pend.end(); ++it) {
            hmpi_request *q = *it;
            if ((q->src == MPI_ANY_SOURCE || q->src == me) && (q->tag == MPI_ANY_TAG || q->tag == tag) && q->epoch == epoch) {
                if (bytes > q->bytes) { fprintf(stderr, "hmpi: message of %zu bytes truncated to %zu\n", bytes, q->bytes); abort(); }
                std::memcpy(q->buf, buf, bytes);
                q->status.MPI_SOURCE = me;
                q->status.MPI_TAG = tag;
                q->complete = true;
                pend.erase(it);
                c->mail_cv.notify_all();
                return;
            }
        }
        Message m;
        m.src = me;
        m.tag = tag;
        m.epoch = epoch;
        m.data.assign((const char *)buf, (const char *)buf + bytes);
        c->mail[(size_t)dst].push_back(std::move(m));
    }
    c->mail_cv.notify_all();
}

void take(hmpi_comm *c, int me, int src, int tag, void *buf, size_t bytes, MPI_Status *st, long epoch = -1) {
    std::unique_lock<std::mutex> lk(c->mail_mu);
    for (;;) {
        auto &q = c->mail[(size_t)me];
        for (auto it = q.begin(); it != q.end(); ++it) {
            if ((src == MPI_ANY_SOURCE || it->src == src) && (tag == MPI_ANY_TAG || it->tag == tag) && it->epoch == epoch) {
                if (it->data.size() > bytes) {
                    fprintf(stderr, "hmpi: message of %zu bytes truncated to %zu\n", it->data.size(), bytes);
                    abort();
                }
                std::memcpy(buf, it->data.data(), it->data.size());
                if (st) { st->MPI_SOURCE = it->src; st->MPI_TAG = it->tag; st->MPI_ERROR = 0; }
                q.erase(it);
                return;
            }
        }
        c->mail_cv.wait(lk);
    }
}

}  // namespace

MPI_Comm hmpi_world() { return g_world; }
int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize() { return MPI_SUCCESS; }
double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int MPI_Comm_rank(MPI_Comm c, int *rank) { *rank = my_rank(c); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { *size = c->size; return MPI_SUCCESS; }

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out) {
    const int me = my_rank(c);
    int mine[2] = {color, key};
    publish(c, me, mine);
    std::vector<std::tuple<int, int, int>> members;  // (key, old rank, world rank)
    int leader = -1;
    for (int i = 0; i < c->size; i++) {
        const int *o = (const int *)c->slot[(size_t)i];
        if (o[0] == color) {
            members.emplace_back(o[1], i, c->world_ranks[(size_t)i]);
            if (leader < 0) leader = i;
        }
    }
    done(c);
    std::stable_sort(members.begin(), members.end());
    hmpi_comm *nc = nullptr;
    if (me == leader) {
        nc = new hmpi_comm((int)members.size());
        for (auto &m : members) nc->world_ranks.push_back(std::get<2>(m));
    }
    publish(c, me, nc);
    nc = (hmpi_comm *)c->slot[(size_t)leader];
    done(c);
    *out = nc;
    return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) { return MPI_Comm_split(c, 0, my_rank(c), out); }
int MPI_Comm_free(MPI_Comm *c) {
    // the last member to arrive deletes; keep it simple: leak-free enough for a test harness
    hmpi_comm *p = *c;
    const int me = my_rank(p);
    p->bar.wait();
    if (me == 0 && p != g_world) delete p;
    *c = nullptr;
    return MPI_SUCCESS;
}
int MPI_Barrier(MPI_Comm c) { c->bar.wait(); return MPI_SUCCESS; }

int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
    const int me = my_rank(c);
    publish(c, me, buf);
    if (me != root) std::memcpy(buf, c->slot[(size_t)root], (size_t)count * type_size(t));
    done(c);
    return MPI_SUCCESS;
}
int MPI_Gather(const void *s, int sc, MPI_Datatype st, void *r, int, MPI_Datatype, int root, MPI_Comm c) {
    const int me = my_rank(c);
    publish(c, me, s);
    const size_t b = (size_t)sc * type_size(st);
    if (me == root)
        for (int i = 0; i < c->size; i++) std::memcpy((char *)r + b * (size_t)i, c->slot[(size_t)i], b);
    done(c);
    return MPI_SUCCESS;
}
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm c) {
    const int me = my_rank(c);
    const size_t b = (size_t)sc * type_size(st);
    if (s == MPI_IN_PLACE) s = (char *)r + b * (size_t)me;
    publish(c, me, s);
    for (int i = 0; i < c->size; i++)
        if ((const char *)r + b * (size_t)i != c->slot[(size_t)i]) std::memcpy((char *)r + b * (size_t)i, c->slot[(size_t)i], b);
    done(c);
    return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *s, int sc, MPI_Datatype st, void *r, const int *rcounts, const int *displs, MPI_Datatype rt,
                   MPI_Comm c) {
    const int me = my_rank(c);
    (void)sc;
    publish(c, me, s);
    const size_t e = type_size(rt);
    (void)st;
    for (int i = 0; i < c->size; i++)
        std::memcpy((char *)r + e * (size_t)displs[i], c->slot[(size_t)i], e * (size_t)rcounts[i]);
    done(c);
    return MPI_SUCCESS;
}
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm c) {
    const int me = my_rank(c);
    const size_t b = (size_t)sc * type_size(st);
    publish(c, me, s);
    for (int i = 0; i < c->size; i++) std::memcpy((char *)r + b * (size_t)i, (const char *)c->slot[(size_t)i] + b * (size_t)me, b);
    done(c);
    return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *s, const int *scounts, const int *sdispls, MPI_Datatype st, void *r, const int *rcounts,
                  const int *rdispls, MPI_Datatype, MPI_Comm c) {
    const int me = my_rank(c);
    const size_t e = type_size(st);
    (void)scounts;
    publish(c, me, s, sdispls);
    for (int i = 0; i < c->size; i++) {
        const int *their_displs = (const int *)c->slot2[(size_t)i];
        std::memcpy((char *)r + e * (size_t)rdispls[i], (const char *)c->slot[(size_t)i] + e * (size_t)their_displs[me],
                    e * (size_t)rcounts[i]);
    }
    done(c);
    return MPI_SUCCESS;
}
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op, MPI_Comm c) {
    const int me = my_rank(c);
    const size_t b = (size_t)count * type_size(t);
    if (s == MPI_IN_PLACE) s = r;
    publish(c, me, s);
    std::vector<char> acc(b);
    std::memcpy(acc.data(), c->slot[0], b);  // rank order 0..n-1
    for (int i = 1; i < c->size; i++) reduce_sum(t, acc.data(), c->slot[(size_t)i], (size_t)count);
    done(c);  // everyone has finished reading everybody's input
    std::memcpy(r, acc.data(), b);
    c->bar.wait();
    return MPI_SUCCESS;
}
int MPI_Reduce_scatter(const void *s, void *r, const int *rcounts, MPI_Datatype t, MPI_Op, MPI_Comm c) {
    const int me = my_rank(c);
    const size_t e = type_size(t);
    size_t off = 0;
    for (int i = 0; i < me; i++) off += (size_t)rcounts[i];
    publish(c, me, s);
    const size_t n = (size_t)rcounts[me];
    std::vector<char> acc(n * e);
    std::memcpy(acc.data(), (const char *)c->slot[0] + off * e, n * e);
    for (int i = 1; i < c->size; i++) reduce_sum(t, acc.data(), (const char *)c->slot[(size_t)i] + off * e, n);
    done(c);
    std::memcpy(r, acc.data(), n * e);
    c->bar.wait();
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
    post(c, my_rank(c), dst, tag, buf, (size_t)count * type_size(t));
    return MPI_SUCCESS;
}
int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
    take(c, my_rank(c), src, tag, buf, (size_t)count * type_size(t), st);
    return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *s, int sc, MPI_Datatype st, int dst, int stag, void *r, int rc, MPI_Datatype rt, int src,
                 int rtag, MPI_Comm c, MPI_Status *status) {
    const int me = my_rank(c);
    long epoch;
    {
        std::lock_guard<std::mutex> lk(c->mail_mu);
        epoch = c->sendrecv_calls[(size_t)me][stag]++;
    }
    post(c, me, dst, stag, s, (size_t)sc * type_size(st), epoch);
    take(c, me, src, rtag, r, (size_t)rc * type_size(rt), status, epoch);
    return MPI_SUCCESS;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *req) {
    MPI_Send(buf, count, t, dst, tag, c);  // eager: complete at once
    *req = new hmpi_request();
    return MPI_SUCCESS;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req) {
    hmpi_request *q = new hmpi_request();
    q->is_recv = true;
    q->buf = buf;
    q->bytes = (size_t)count * type_size(t);
    q->src = src;
    q->tag = tag;
    q->comm = c;
    const int me = my_rank(c);
    {
        std::lock_guard<std::mutex> lk(c->mail_mu);
        auto &box = c->mail[(size_t)me];
        bool matched = false;
        for (auto it = box.begin(); it != box.end(); ++it) {
            if ((src == MPI_ANY_SOURCE || it->src == src) && (tag == MPI_ANY_TAG || it->tag == tag) && it->epoch == -1) {
                if (it->data.size() > q->bytes) { fprintf(stderr, "hmpi: message truncated\n"); abort(); }
                std::memcpy(buf, it->data.data(), it->data.size());
                q->status.MPI_SOURCE = it->src;
                q->status.MPI_TAG = it->tag;
                q->complete = true;
                box.erase(it);
                matched = true;
                break;
            }
        }
        if (!matched) c->pending[(size_t)me].push_back(q);
    }
    *req = q;
    return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *st) {
    hmpi_request *q = *req;
    if (q && q->is_recv) {
        std::unique_lock<std::mutex> lk(q->comm->mail_mu);
        q->comm->mail_cv.wait(lk, [&] { return q->complete; });
        if (st) *st = q->status;
    }
    delete q;
    *req = nullptr;
    return MPI_SUCCESS;
}

int MPI_Type_create_struct(int n, const int *blocklens, const MPI_Aint *offsets, const MPI_Datatype *types, MPI_Datatype *out) {
    size_t extent = 0;
    for (int i = 0; i < n; i++) extent = std::max(extent, (size_t)offsets[i] + (size_t)blocklens[i] * type_size(types[i]));
    extent = (extent + 7) & ~(size_t)7;  // alignment of the members used here (8-byte)
    std::lock_guard<std::mutex> lk(g_type_mu);
    g_struct_sizes.push_back(extent);
    *out = 100 + (int)g_struct_sizes.size() - 1;
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }

void hmpi_run(int p, int threads_per_rank, void (*fn)(int, void *), void *arg) {
    hmpi_comm *world = new hmpi_comm(p);
    for (int i = 0; i < p; i++) world->world_ranks.push_back(i);
    g_world = world;
    std::vector<std::thread> ts;
    for (int r = 0; r < p; r++)
        ts.emplace_back([=] {
            tl_world_rank = r;
            if (threads_per_rank > 0) omp_set_num_threads(threads_per_rank);
            fn(r, arg);
        });
    for (auto &t : ts) t.join();
    g_world = nullptr;
    delete world;
}
