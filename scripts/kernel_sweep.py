#!/usr/bin/env python
"""Per-kernel roofline sweep over the block shapes of BASELINE.json's configs (one GPU).
Writes gpurun_out/kernel_sweep.json; summarised into profiles/ by hand."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes, measured_peaks  # noqa: E402
from distributed_sddmm_b200 import check, lib  # noqa: E402


def generate_er(L, logM, npr, seed, row_lo, row_hi):
    cap = (row_hi - row_lo) * npr
    r = np.empty(cap, np.uint64)
    c = np.empty(cap, np.uint64)
    v = np.empty(cap, np.float64)
    n = L.hnh_er_generate_host(logM, npr, seed, row_lo, row_hi, r.ctypes.data, c.ctypes.data, v.ctypes.data, cap)
    assert n >= 0
    return r[:n], c[:n], v[:n]


def build_csr(L, rows, cols, r, c, v):
    nnz = len(r)
    rs = np.empty(rows + 1, np.int64)
    ci = np.empty(max(nnz, 1), np.int64)
    ri = np.empty(max(nnz, 1), np.int64)
    vv = np.empty(max(nnz, 1), np.float64)
    rc = L.hnh_coo_to_csr_host(rows, cols, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data, 0, rs.ctypes.data,
                               ci.ctypes.data, ri.ctypes.data, vv.ctypes.data)
    assert rc == 0
    return rs, ci[:nnz], ri[:nnz], vv[:nnz]


L = lib()
dev = torch.device("cuda:0")
peak, kind = measured_peaks()
BETA0 = 4
flags_extra = int(os.environ.get("HNH_SWEEP_FLAGS", "0"))

# (label, logN (columns), rows, nnz/row, r)
SHAPES = [
    ("cfg2@1 r128", 20, 1 << 20, 32, 128),
    ("cfg4-like r256", 20, 1 << 20, 32, 256),
    ("r64", 20, 1 << 20, 32, 64),
    ("r32", 20, 1 << 20, 32, 32),
    ("cfg3 c=8 r32", 22, 1 << 19, 64, 32),
    ("cfg3 c=4 r16", 22, 1 << 19, 64, 16),
    ("cfg3 c=2 r8", 22, 1 << 19, 64, 8),
    ("cfg3 c=1 r4", 22, 1 << 19, 64, 4),
    ("cfg2@8 block r128", 17, 1 << 17, 4, 128),
    ("cfg5@8 block r128", 18, 1 << 18, 4, 128),
]
if len(sys.argv) > 1:
    SHAPES = [s for s in SHAPES if any(a in s[0] for a in sys.argv[1:])]

out = []
for label, logN, rows, npr, r in SHAPES:
    N = 1 << logN
    rr, cc, vv = generate_er(L, logN, npr, 0xC0FFEE + 9, 0, rows)
    rs, ci, ri, vals = build_csr(L, rows, N, rr, cc, vv)
    nnz = len(ci)
    d_rs, d_ci, d_ri = (torch.from_numpy(x).to(dev) for x in (rs, ci, ri))
    d_v = torch.zeros(nnz, dtype=torch.float64, device=dev)
    X = torch.full((rows, r), 0.001, dtype=torch.float64, device=dev)
    Y = torch.full((N, r), 0.001, dtype=torch.float64, device=dev)
    O = torch.zeros((rows, r), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def t(fn, n=5):
        for _ in range(2):
            fn()
        ms = []
        for _ in range(n):
            flush.zero_()  # L2 flush between timed launches (small blocks fit in L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    P = lambda x: x.data_ptr()
    kern = {
        "sddmm": (lambda: check(L.hnh_sddmm_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(X), P(Y), r, flags_extra, st)), False),
        "sddmm_b0": (lambda: check(L.hnh_sddmm_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(X), P(Y), r, BETA0 | flags_extra, st)), True),
        "sddmm_coo": (lambda: check(L.hnh_sddmm_coo_f64(P(d_ri), P(d_ci), P(d_v), nnz, P(X), P(Y), r, flags_extra, st)), False),
        "spmm": (lambda: check(L.hnh_spmm_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(Y), P(O), r, flags_extra, st)), False),
        "spmm_b0": (lambda: check(L.hnh_spmm_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(Y), P(O), r, BETA0 | flags_extra, st)), True),
        "fused": (lambda: check(L.hnh_fused_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(X), P(Y), P(O), r, flags_extra, st)), False),
        "fused_b0": (lambda: check(L.hnh_fused_f64(P(d_rs), P(d_ci), P(d_v), rows, nnz, P(X), P(Y), P(O), r, BETA0 | flags_extra, st)), True),
    }
    for name, (fn, b0) in kern.items():
        base = name.split("_")[0]
        ms = t(fn)
        by = algorithmic_bytes(base, nnz, rows, r, beta0=b0)
        if name == "sddmm_coo":
            by += (nnz - rows - 1) * 8  # row_idx instead of rowStart
        gbs = by / ms / 1e6
        rec = {"shape": label, "rows": rows, "cols": N, "nnz": nnz, "r": r, "kernel": name, "ms": ms,
               "alg_GB": by / 1e9, "GBps": gbs, "frac": gbs / peak, "peak_kind": kind,
               "gflops": (4.0 if base == "fused" else 2.0) * nnz * r / ms / 1e6}
        out.append(rec)
        print(f"{label:20s} {name:10s} {ms:8.3f} ms  {gbs:8.1f} GB/s  frac {gbs/peak:5.3f}", flush=True)
    del d_rs, d_ci, d_ri, d_v, X, Y, O, flush
    torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("HNH_SWEEP_OUT", "kernel_sweep.json")), "w"), indent=1)
