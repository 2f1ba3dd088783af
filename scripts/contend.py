"""Single-process 2-GPU experiment: how fast is a 512 MiB peer copy (copy engine,
cudaMemcpyPeerAsync through torch) while the fused kernel saturates HBM on both GPUs, compared with
the copy alone?  Decides whether the ring shift should use DMA instead of NCCL's SM copy kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from distributed_sddmm_b200 import check, lib

L = lib()
logN, rows, npr, r = 19, 1 << 19, 32, 128
N = 1 << logN


def setup(dev):
    torch.cuda.set_device(dev)
    cap = rows * npr
    rr = np.empty(cap, np.uint64); cc = np.empty(cap, np.uint64); vv = np.empty(cap, np.float64)
    n = L.hnh_er_generate_host(logN, npr, 7, 0, rows, rr.ctypes.data, cc.ctypes.data, vv.ctypes.data, cap)
    rs = np.empty(rows + 1, np.int64); ci = np.empty(n, np.int64); ri = np.empty(n, np.int64); va = np.empty(n, np.float64)
    check(L.hnh_coo_to_csr_host(rows, N, n, rr.ctypes.data, cc.ctypes.data, vv.ctypes.data, 0, rs.ctypes.data, ci.ctypes.data, ri.ctypes.data, va.ctypes.data))
    d = torch.device("cuda", dev)
    return dict(rs=torch.from_numpy(rs).to(d), ci=torch.from_numpy(ci).to(d), v=torch.zeros(n, dtype=torch.float64, device=d),
                X=torch.full((rows, r), 0.001, dtype=torch.float64, device=d), Y=torch.full((N, r), 0.001, dtype=torch.float64, device=d),
                O=torch.zeros((rows, r), dtype=torch.float64, device=d), n=n, stream=torch.cuda.Stream(d), cstream=torch.cuda.Stream(d),
                send=torch.full((N, r), 1.0, dtype=torch.float64, device=d), recv=torch.empty((N, r), dtype=torch.float64, device=d))


g = [setup(0), setup(1)]


def kernel(dev):
    s = g[dev]
    torch.cuda.set_device(dev)
    check(L.hnh_fused_f64(s["rs"].data_ptr(), s["ci"].data_ptr(), s["v"].data_ptr(), rows, s["n"], s["X"].data_ptr(), s["Y"].data_ptr(),
                          s["O"].data_ptr(), r, 48, s["stream"].cuda_stream))


def copy(dev):
    s, o = g[dev], g[1 - dev]
    with torch.cuda.stream(s["cstream"]):
        o["recv"].copy_(s["send"], non_blocking=True)  # push to the other GPU


def sync():
    for d in (0, 1):
        torch.cuda.synchronize(d)


def timed(fn_k, fn_c, reps=5):
    out = []
    for _ in range(reps):
        sync()
        ev = []
        for d in (0, 1):
            torch.cuda.set_device(d)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if fn_c:
                e[0].record(g[d]["cstream"])
            if fn_k:
                e[2].record(g[d]["stream"])
            ev.append(e)
        for d in (0, 1):
            if fn_c:
                fn_c(d)
            if fn_k:
                fn_k(d); fn_k(d)
        for d in (0, 1):
            torch.cuda.set_device(d)
            if fn_c:
                ev[d][1].record(g[d]["cstream"])
            if fn_k:
                ev[d][3].record(g[d]["stream"])
        sync()
        out.append((ev[0][0].elapsed_time(ev[0][1]) if fn_c else 0.0, ev[0][2].elapsed_time(ev[0][3]) if fn_k else 0.0))
    return np.median(np.array(out), axis=0)


mb = N * r * 8 / 2 ** 20
for name, k, c in (("copy alone", None, copy), ("kernel x2 alone", kernel, None), ("copy + kernel x2", kernel, copy)):
    t = timed(k, c)
    print(f"{name:18s} copy {t[0]:7.3f} ms ({mb / 1024 / max(t[0], 1e-9) * 1000:7.1f} GiB/s)   kernels {t[1]:7.3f} ms")
