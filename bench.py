#!/usr/bin/env python
"""bench.py -- FusedMM (SDDMM+SpMM) throughput of the B200-native HnH engine.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 launched under
torch.distributed.run, one rank per GPU) prints ONE JSON line on rank 0.

* workload  : BASELINE.json configs[1] -- Erdos-Renyi N=2^20, 32 nnz/row, r=128, 1.5D
              dense-shift FusedMM (`fusedSpMM(A, B, S, result, Amat)`,
              reference benchmark_dist.cpp:117-141; A=B=0.001, S=1.0, :102-106).
* step      : one fusedSpMM call over the whole distributed matrix.
* metric    : SDDMM+SpMM GFLOP/s with the reference's FLOP model 2*nnz*2*R per FusedMM
              (benchmark_dist.cpp:147).
* value     : inputs resident in HBM.   e2e: host buffers, H2D/D2H inside the timed region.
* roofline  : dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json.
* cpu_baseline / --impl reference : the CPU oracle (restatement of the reference's
              OpenMP SDDMM loop + CSR SpMM; MKL/MPI are not in the image) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0xC0FFEE + 2  # config index 2
FALLBACK_HBM_GBS = 6650.0


# ------------------------------------------------------------------ helpers ---------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def algorithmic_bytes(kind: str, nnz: int, m: int, r: int, beta0: bool = False) -> int:
    """fp64 w=8, int64 x=8.  beta0 variants do not read the old output."""
    w = x = 8
    if kind == "sddmm":
        return nnz * (r * w + x + (w if beta0 else 2 * w)) + (m + 1) * x + m * r * w
    if kind == "spmm":
        return nnz * (r * w + x + w) + (m + 1) * x + (1 if beta0 else 2) * m * r * w
    if kind == "fused":
        return nnz * (r * w + x + (w if beta0 else 2 * w)) + (m + 1) * x + (2 if beta0 else 3) * m * r * w
    raise ValueError(kind)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def generate_er(L, logM, npr, seed, row_lo, row_hi):
    cap = (row_hi - row_lo) * npr
    r = np.empty(cap, np.uint64)
    c = np.empty(cap, np.uint64)
    v = np.empty(cap, np.float64)
    n = L.hnh_er_generate_host(logM, npr, seed, row_lo, row_hi, r.ctypes.data, c.ctypes.data,
                               v.ctypes.data, cap)
    if n < 0:
        raise RuntimeError(L.hnh_last_error_string().decode())
    return r[:n], c[:n], v[:n]


def build_csr(L, rows, cols, r, c, v, transpose=False):
    nnz = len(r)
    out_rows = cols if transpose else rows
    rs = np.empty(out_rows + 1, np.int64)
    ci = np.empty(max(nnz, 1), np.int64)
    ri = np.empty(max(nnz, 1), np.int64)
    vv = np.empty(max(nnz, 1), np.float64)
    rc = L.hnh_coo_to_csr_host(rows, cols, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data,
                               int(transpose), rs.ctypes.data, ci.ctypes.data, ri.ctypes.data,
                               vv.ctypes.data)
    if rc:
        raise RuntimeError(L.hnh_last_error_string().decode())
    return rs, ci[:nnz], ri[:nnz], vv[:nnz]


# ------------------------------------------------------------------ CPU arm ---------------
def cpu_fusedmm_sample(L, args, budget_s=12.0):
    """Time the CPU oracle (reference restatement) on a bounded row-sample of the workload:
    the first m_s rows of S against the full B -- per-row work is identical to the full job,
    so GFLOP/s carries over.  Returns (gflops, cores, description)."""
    from oracle import hnh_oracle as orc
    N = 1 << args.logM
    R = args.R
    rng_rows = min(N, 1 << 16)
    r, c, v = generate_er(L, args.logM, args.nnz_per_row, SEED, 0, rng_rows)
    A = np.full((N, R), 0.001)
    B = np.full((N, R), 0.001)

    def run(mrows, trials):
        rr, cc, vv = (r, c, v) if mrows == rng_rows else generate_er(L, args.logM, args.nnz_per_row, SEED, 0, mrows)
        rs, ci, ri, _ = build_csr(L, mrows, N, rr, cc, vv)
        vals = np.zeros(len(ci))
        out = np.zeros((mrows, R))
        orc.fused_block(rs, ri, ci, vals, A[:mrows], B, out)  # warm-up
        t0 = time.perf_counter()
        for _ in range(trials):
            vals[:] = 0.0
            out[:] = 0.0
            orc.fused_block(rs, ri, ci, vals, A[:mrows], B, out)
        dt = (time.perf_counter() - t0) / trials
        return 4.0 * len(ci) * R / dt / 1e9, dt, len(ci)

    g, dt, nnz = run(rng_rows, 1)
    # scale the sample so that 3 trials take about budget_s
    scale = max(1.0, min(N / rng_rows, budget_s / 3.0 / max(dt, 1e-6)))
    mrows = int(min(N, rng_rows * scale))
    mrows = max(rng_rows, (mrows >> 12) << 12)
    if mrows > rng_rows:
        g, dt, nnz = run(mrows, 3)
    cores = orc.num_threads()
    desc = (f"CPU restatement of reference (MKL/MPI unavailable in image): OpenMP SDDMM loop + CSR "
            f"SpMM on the first {mrows} of {N} rows (nnz={nnz}) against the full B, "
            f"{cores} threads, {dt*1e3:.1f} ms per FusedMM sample")
    return g, cores, desc, dt, mrows, nnz


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from distributed_sddmm_b200 import lib
    L = lib()
    # K steps of a bounded sample each; keep the whole run within a few minutes
    per_step = max(1.0, min(12.0, 120.0 / max(1, args.steps + args.warmup)))
    samples = []
    desc = cores = None
    for i in range(args.warmup + args.steps):
        g, cores, desc, dt, mrows, nnz = cpu_fusedmm_sample(L, args, budget_s=per_step * 3)
        if i >= args.warmup:
            samples.append((g, dt))
    g = float(np.mean([s[0] for s in samples]))
    ms = float(np.mean([s[1] for s in samples])) * 1e3
    line = {
        "impl": "reference", "metric": "SDDMM+SpMM GFLOP/s (FusedMM, 4*nnz*R flop)", "value": g,
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, "cpu"),
        "cpu_baseline": {"value": g, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": g, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(args, where):
    return {"workload": f"Erdos-Renyi N=2^{args.logM} nnz/row={args.nnz_per_row} r={args.R} FusedMM "
                        f"(fusedSpMM, 1.5D dense-shift, {args.alg})",
            "logM": args.logM, "nnz_per_row": args.nnz_per_row, "R": args.R, "algorithm": args.alg,
            "c": args.c, "seed": SEED, "index_type": "int64", "where": where,
            "l2": "inputs (dense factors + CSR) are far larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------ GPU arm ---------------
def run_native(args):
    import torch
    from distributed_sddmm_b200 import lib, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    L = lib()
    if world > 1:
        from distributed_sddmm_b200 import dist_bench
        return dist_bench.run(args)

    dev = torch.device("cuda", local_rank)
    N, R = 1 << args.logM, args.R
    r, c, v = generate_er(L, args.logM, args.nnz_per_row, SEED, 0, N)
    rs, ci, ri, vals = build_csr(L, N, N, r, c, v)
    nnz = len(ci)
    del r, c, v
    d_rs, d_ci = torch.from_numpy(rs).to(dev), torch.from_numpy(ci).to(dev)
    d_vals = torch.zeros(nnz, dtype=torch.float64, device=dev)
    d_S = torch.ones(nnz, dtype=torch.float64, device=dev)       # like_S_values(1.0)
    d_res = torch.zeros(nnz, dtype=torch.float64, device=dev)    # sddmm_result
    A = torch.full((N, R), 0.001, dtype=torch.float64, device=dev)  # like_A_matrix(0.001)
    B = torch.full((N, R), 0.001, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    BETA0 = 4
    kev = []  # (start, end) events around the dominant kernel

    def step_fusion2(record=False):
        # Sparse15D_Dense_Shift::fusedSpMM, fusion 2, p = c = 1 (15D_dense_shift.hpp:151-252):
        # values = 0; K1; K2 into the accumulator; A = accumulator -- one fused kernel, in place.
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(L.hnh_fused_f64(d_rs.data_ptr(), d_ci.data_ptr(), d_vals.data_ptr(), N, nnz,
                              A.data_ptr(), B.data_ptr(), A.data_ptr(), R, BETA0, st), "fused")
        if record:
            e1.record()
            kev.append((e0, e1))

    def step_fusion1(record=False):
        # Distributed_Sparse::fusedSpMM (distributed_sparse.h:296-312) = sddmmA, A.setZero(),
        # spmmA with the SDDMM result, p = c = 1.
        check(L.hnh_sddmm_f64(d_rs.data_ptr(), d_ci.data_ptr(), d_vals.data_ptr(), N, nnz,
                              A.data_ptr(), B.data_ptr(), R, BETA0, st), "sddmm")
        check(L.hnh_hadamard_f64(d_res.data_ptr(), d_S.data_ptr(), d_vals.data_ptr(), nnz, st), "hadamard")
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(L.hnh_spmm_f64(d_rs.data_ptr(), d_ci.data_ptr(), d_res.data_ptr(), N, nnz,
                             B.data_ptr(), A.data_ptr(), R, BETA0, st), "spmm")
        if record:
            e1.record()
            kev.append((e0, e1))

    step = step_fusion2 if args.alg == "15d_fusion2" else step_fusion1
    dominant = "fused" if args.alg == "15d_fusion2" else "spmm"

    def reset_inputs():
        A.fill_(0.001)

    def timed(fn, steps, warmup, record):
        for _ in range(warmup):
            reset_inputs()
            fn(False)
        torch.cuda.synchronize()
        t = 0.0
        for _ in range(steps):
            reset_inputs()  # untimed: restore the benchmark inputs (A is overwritten by FusedMM)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(record)
            e1.record()
            torch.cuda.synchronize()
            t += e0.elapsed_time(e1)
        return t / steps

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.hnh_launch_count()
    ms = timed(step, args.steps, args.warmup, True)
    launches = (L.hnh_launch_count() - launches0)
    clocks = sampler.stop()
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    flops = 4.0 * nnz * R
    gflops = flops / (ms * 1e-3) / 1e9

    # dominant-kernel roofline
    kms = float(np.mean([a.elapsed_time(b) for a, b in kev])) if kev else ms
    bytes_alg = algorithmic_bytes(dominant, nnz, N, R, beta0=True)
    peak, peak_kind = measured_peaks()
    achieved = bytes_alg / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_kind": f"of {peak_kind}",
                "kernel": f"{dominant}_row_kernel<{R}> (BETA0)", "kernel_ms": kms,
                "algorithmic_bytes_per_launch": bytes_alg}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(f"{dominant}_{R}")
        except Exception:
            pass

    # the other variant, for the record (not the headline)
    other_name = "15d_fusion1" if args.alg == "15d_fusion2" else "15d_fusion2"
    other = step_fusion1 if args.alg == "15d_fusion2" else step_fusion2
    oms = timed(other, max(3, args.steps // 2), 2, False)

    # ---- e2e: host buffers through the C-ABI block API, copies inside the timed region ----
    import ctypes as C
    blk = C.c_void_p()
    check(L.hnh_block_create_host(rs.ctypes.data, ci.ctypes.data, N, N, nnz, R, C.byref(blk)), "block_create")
    hA = torch.full((N, R), 0.001, dtype=torch.float64).pin_memory()
    hB = torch.full((N, R), 0.001, dtype=torch.float64).pin_memory()
    hV = torch.zeros(nnz, dtype=torch.float64).pin_memory()
    hO = torch.zeros((N, R), dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        check(L.hnh_block_run_host(blk, 2, hA.data_ptr(), hB.data_ptr(), hV.data_ptr(), hO.data_ptr(),
                                   R, 3, st), "block_run_host")

    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    L.hnh_block_destroy(blk)
    h2d = 2 * N * R * 8
    d2h = N * R * 8 + nnz * 8
    e2e = {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "hnh_block_run_host(op=fused): pinned host A,B in; A result + SDDMM values out"}
    del hA, hB, hV, hO

    # ---- CPU baseline on the host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu_baseline:
        g, cores, desc, _, _, _ = cpu_fusedmm_sample(L, args)
        cpu = {"value": g, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": desc}

    line = {
        "metric": "SDDMM+SpMM GFLOP/s (FusedMM, 4*nnz*R flop)", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(workload_config(args, "cuda"), nnz=nnz, p=1),
        "hbm_gbs_achieved": (bytes_alg if args.alg == "15d_fusion2" else
                             algorithmic_bytes("sddmm", nnz, N, R, True) + 3 * 8 * nnz +
                             algorithmic_bytes("spmm", nnz, N, R, True)) / (ms * 1e-3) / 1e9,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_timed),
        "clocks": clocks,
        "other": {other_name: {"ms_per_step": oms, "gflops": flops / (oms * 1e-3) / 1e9}},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--logM", type=int, default=20)
    ap.add_argument("--nnz-per-row", type=int, default=32)
    ap.add_argument("--R", type=int, default=128)
    ap.add_argument("--c", type=int, default=1)
    ap.add_argument("--alg", default="15d_fusion2", choices=["15d_fusion1", "15d_fusion2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
