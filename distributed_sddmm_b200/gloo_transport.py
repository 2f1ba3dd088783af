"""The External transport of hnh/comm.h implemented with torch.distributed (gloo) callbacks.

Used by the CPU tests (world_size 2, no GPU: the host-side setup path of every algorithm) and
by the single-GPU multi-process tests (several ranks sharing cuda:0, device buffers staged through
pinned host memory by the C++ side).  A maintainer with a real MPI would fill the same callback
table with MPI_Sendrecv / MPI_Allgather / ...."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as B


def _view(ptr, nbytes):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)


class GlooTransport:
    def __init__(self):
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        self.groups = {0: (None, list(range(dist.get_world_size())))}  # id -> (group, world ranks)
        self.next_id = 1
        self.cb = dict(
            sendrecv=B.CB_SENDRECV(self._sendrecv), allgather=B.CB_ALLGATHER(self._allgather),
            reduce_scatter_f64=B.CB_REDUCE_SCATTER(self._reduce_scatter), allreduce_f64=B.CB_ALLREDUCE(self._allreduce),
            alltoallv=B.CB_ALLTOALLV(self._alltoallv), barrier=B.CB_BARRIER(self._barrier), split=B.CB_SPLIT(self._split))
        self.table = B.ExternalTransport(None, *[self.cb[k] for k in ("sendrecv", "allgather", "reduce_scatter_f64",
                                                                      "allreduce_f64", "alltoallv", "barrier", "split")])

    def _guard(self, fn):
        try:
            fn()
            return 0
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            return 1

    def _sendrecv(self, ctx, comm, sbuf, sbytes, dst, rbuf, rbytes, src):
        def run():
            g, ranks = self.groups[comm]
            me = ranks.index(dist.get_rank())
            s, r = _view(sbuf, sbytes), _view(rbuf, rbytes)
            if dst == me and src == me:
                r.copy_(s)
                return
            reqs = []
            if sbytes:
                reqs.append(dist.isend(s.clone(), ranks[dst], group=g))
            if rbytes:
                reqs.append(dist.irecv(r, ranks[src], group=g))
            for q in reqs:
                q.wait()
        return self._guard(run)

    def _allgather(self, ctx, comm, sbuf, rbuf, nbytes):
        def run():
            g, ranks = self.groups[comm]
            s = _view(sbuf, nbytes).clone()
            out = _view(rbuf, nbytes * len(ranks))
            if len(ranks) == 1:
                out.copy_(s)
            else:
                self._ag_list(out, s, g, len(ranks), nbytes)
        return self._guard(run)

    @staticmethod
    def _ag_list(out, s, g, n, nbytes):
        parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(n)]
        dist.all_gather(parts, s, group=g)
        for i, p in enumerate(parts):
            out[i * nbytes:(i + 1) * nbytes].copy_(p)

    def _reduce_scatter(self, ctx, comm, sbuf, rbuf, count):
        def run():
            g, ranks = self.groups[comm]
            n = len(ranks)
            full = _view(sbuf, count * n * 8).view(torch.float64).clone()
            if n > 1:
                dist.all_reduce(full, group=g)
            me = ranks.index(dist.get_rank())
            _view(rbuf, count * 8).view(torch.float64).copy_(full[me * count:(me + 1) * count])
        return self._guard(run)

    def _allreduce(self, ctx, comm, buf, count):
        def run():
            g, ranks = self.groups[comm]
            t = _view(buf, count * 8).view(torch.float64)
            if len(ranks) > 1 and count:
                tmp = t.clone()
                dist.all_reduce(tmp, group=g)
                t.copy_(tmp)
        return self._guard(run)

    def _alltoallv(self, ctx, comm, sbuf, sbytes, sdispls, rbuf, rbytes, rdispls):
        def run():
            g, ranks = self.groups[comm]
            n = len(ranks)
            me = ranks.index(dist.get_rank())
            sb = [sbytes[i] for i in range(n)]
            sd = [sdispls[i] for i in range(n)]
            rb = [rbytes[i] for i in range(n)]
            rd = [rdispls[i] for i in range(n)]
            s = _view(sbuf, max([a + b for a, b in zip(sd, sb)] + [0]))
            r = _view(rbuf, max([a + b for a, b in zip(rd, rb)] + [0]))
            reqs = []
            for i in range(n):
                if i == me:
                    if sb[i]:
                        r[rd[i]:rd[i] + rb[i]].copy_(s[sd[i]:sd[i] + sb[i]])
                    continue
                if sb[i]:
                    reqs.append(dist.isend(s[sd[i]:sd[i] + sb[i]].clone(), ranks[i], group=g))
            for i in range(n):
                if i != me and rb[i]:
                    tmp = torch.empty(rb[i], dtype=torch.uint8)
                    dist.recv(tmp, ranks[i], group=g)
                    r[rd[i]:rd[i] + rb[i]].copy_(tmp)
            for q in reqs:
                q.wait()
        return self._guard(run)

    def _barrier(self, ctx, comm):
        def run():
            g, ranks = self.groups[comm]
            if len(ranks) > 1:
                dist.barrier(group=g)
        return self._guard(run)

    def _split(self, ctx, comm, color, key, new_comm, new_rank, new_size):
        def run():
            g, ranks = self.groups[comm]
            me = dist.get_rank()
            mine = torch.tensor([color, key, me], dtype=torch.int64)
            allv = [torch.zeros(3, dtype=torch.int64) for _ in ranks]
            if len(ranks) > 1:
                dist.all_gather(allv, mine, group=g)
            else:
                allv = [mine]
            colors = sorted({int(t[0]) for t in allv})
            my_group, my_ranks = None, None
            for col in colors:  # every member creates every group, in the same order
                members = sorted([(int(t[1]), int(t[2])) for t in allv if int(t[0]) == col])
                wr = [m[1] for m in members]
                grp = dist.new_group(wr) if len(wr) > 1 else None
                if col == color:
                    my_group, my_ranks = grp, wr
            cid = self.next_id
            self.next_id += 1
            self.groups[cid] = (my_group, my_ranks)
            new_comm[0] = cid
            new_rank[0] = my_ranks.index(me)
            new_size[0] = len(my_ranks)
        return self._guard(run)
