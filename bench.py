#!/usr/bin/env python
"""bench.py -- FusedMM (SDDMM+SpMM) throughput of the B200-native HnH engine.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` (N>1 launched under
torch.distributed.run, one rank per GPU) prints ONE JSON line on rank 0.

* workload  : BASELINE.json configs[1] -- Erdos-Renyi N=2^20, 32 nnz/row, r=128, 1.5D
              dense-shift FusedMM: `fusedSpMM(A, B, S, result, Amat)` of
              Sparse15D_Dense_Shift, exactly what bench_erdos_renyi / benchmark_algorithm time
              (reference benchmark_dist.cpp:117-141; A=B=0.001, S=1.0, :102-106).  The C++ host
              classes of libhnh_b200.so run it; this file only drives them (ctypes) and measures.
* step      : one fusedSpMM call over the whole distributed matrix (strong scaling: the matrix
              is fixed, N GPUs share it).
* metric    : SDDMM+SpMM GFLOP/s with the reference's FLOP model 2*nnz*2*R per FusedMM
              (benchmark_dist.cpp:147).
* value     : inputs resident in HBM.   e2e: per-rank pinned host buffers, H2D/D2H inside the
              timed region.
* roofline  : local kernels' algorithmic bytes / their CUDA-event time vs MEASURED_PEAKS.json.
* cpu_baseline / --impl reference : the REFERENCE's own code (oracle/_ref: its sources compiled
              unmodified against MPI/MKL/Eigen/CombBLAS shims) on the host cores; falls back to
              the C port of its kernels (oracle/hnh_oracle.c) where oracle/_ref is not built.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0xC0FFEE + 2  # config index 2
FALLBACK_HBM_GBS = 6650.0
METRIC = "SDDMM+SpMM GFLOP/s (FusedMM, 4*nnz*R flop)"


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
    """fp64 w=8, int64 x=8, one local kernel call on a block with nnz nonzeros and
    m CSR rows.  beta0 variants do not read the old output / values."""
    w = x = 8
    if kind == "sddmm":
        return nnz * (r * w + x + (w if beta0 else 2 * w)) + (m + 1) * x + m * r * w
    if kind == "spmm":
        return nnz * (r * w + x + w) + (m + 1) * x + (1 if beta0 else 2) * m * r * w
    if kind == "fused":
        return nnz * (r * w + x + (w if beta0 else 2 * w)) + (m + 1) * x + (2 if beta0 else 3) * m * r * w
    raise ValueError(kind)


def fusedmm_bytes_per_rank(alg: str, nnz_rank: float, rows_stationary: int, steps: int, r: int) -> float:
    """Algorithmic HBM bytes one rank's local kernels move in one FusedMM (sum over ring steps).
    fusion 2: `steps` fused kernels on blocks of nnz_rank/steps nonzeros and rows_stationary rows;
    values written once, the accumulator written at step 0 and read+written afterwards.
    fusion 1: an SDDMM pass and an SpMM pass over the same blocks + the Hadamard / setCSRValues
    plumbing (3 + 2 value passes)."""
    w = x = 8
    m = rows_stationary
    if alg == "15d_fusion2":
        return nnz_rank * (r * w + x + w) + steps * ((m + 1) * x + m * r * w) + (2 * steps - 1) * m * r * w
    sddmm = nnz_rank * (r * w + x + w) + steps * ((m + 1) * x + m * r * w)
    spmm = nnz_rank * (r * w + x + w) + steps * ((m + 1) * x + 2 * m * r * w)
    return sddmm + spmm + 5 * w * nnz_rank


NVLINK_GBS_NOMINAL = 900.0  # NVLink 5, per GPU and direction (B200_PROFILING.md); MEASURED_PEAKS.json has no link figure


def nvlink_bytes_per_rank(alg: str, p: int, c: int, local_rows: int, r: int) -> float:
    """Bytes one rank RECEIVES over NVLink in one FusedMM of the 1.5D dense-shift algorithm:
    every ring pass delivers p/c - 1 riding shards of local_rows x r doubles; replication (c > 1) adds the
    all-gather of the stationary operand (c - 1 shards in) and, for local kernel fusion, the reduce-scatter of
    the c x larger accumulator (c - 1 shard-sized partial sums in).  Fusion 2 makes one ring pass per FusedMM.
    Fusion 1 (replication reuse) makes two -- the SDDMM pass, and the SpMM pass in which the OUTPUT rides and
    therefore goes all the way round (p/c shifts) -- gathers once and never reduce-scatters."""
    shard = 8.0 * local_rows * r
    steps = p // c
    gather = (c - 1) * shard
    if alg == "15d_fusion2":
        return (steps - 1) * shard + 2 * gather
    return ((steps - 1) + (steps if steps > 1 else 0)) * shard + gather


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (NVML, 5 ms period)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"nvml unavailable: {e}"
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "no samples"]}
        self._stop.set()
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": sorted(self.reasons)}


def workload_config(args, where, **extra):
    cfg = {"workload": f"Erdos-Renyi N=2^{args.logM} nnz/row={args.nnz_per_row} r={args.R} FusedMM "
                       f"(Sparse15D_Dense_Shift::fusedSpMM, {args.alg})",
           "logM": args.logM, "nnz_per_row": args.nnz_per_row, "R": args.R, "algorithm": args.alg, "c": args.c,
           "seed": SEED, "index_type": "int64", "where": where,
           "l2": "per-GPU working set (dense shards + CSR) exceeds the 126 MB L2 and every step gathers "
                 "from a different block; no explicit flush between steps"}
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------ CPU arm ---------------
class _quiet_stdout:
    """The reference prints progress with cout; keep bench.py's stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def cpu_reference_fusedmm(args, warmup, steps, budget_s=45.0):
    """GFLOP/s of the reference's fusedSpMM on the host cores (rank 0 only).  Uses oracle/_ref
    (the reference's own code) when built, else the C port of its kernels.  The sample is the
    whole configured matrix when that fits the time budget, otherwise the same configuration at
    a smaller logM (same nnz/row and r: identical per-nonzero work)."""
    from oracle import ref
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    if ref.available():
        with _quiet_stdout():
            nnz0, s0 = ref.time_fused(args.alg, 1, 1, args.R, min(args.logM, 16), args.nnz_per_row, SEED, 1, 1, cores)
        gf0 = 4.0 * nnz0 * args.R / s0[-1] / 1e9
        logM = args.logM
        while logM > 16:
            flop = 4.0 * (1 << logM) * args.nnz_per_row * args.R
            # fusedSpMM calls + construction (redistribution, sorts, two COO->CSR), ~6 call-equivalents
            if (warmup + steps + 6) * flop / (gf0 * 1e9) <= budget_s:
                break
            logM -= 1
        with _quiet_stdout():
            nnz, secs = ref.time_fused(args.alg, 1, 1, args.R, logM, args.nnz_per_row, SEED, warmup, steps, cores)
        per = secs[warmup:]
        gf = 4.0 * nnz * args.R / per / 1e9
        desc = (f"the reference's own code (oracle/_ref: reference sources compiled unmodified; MKL/MPI/Eigen/"
                f"CombBLAS shimmed) {args.alg} p=1, {cores} OpenMP threads, Erdos-Renyi N=2^{logM} "
                f"nnz/row={args.nnz_per_row} r={args.R} (nnz={nnz}), {len(per)} fusedSpMM calls after {warmup} warm-up, "
                f"{np.mean(per)*1e3:.1f} ms each")
        return float(np.mean(gf)), cores, "reference", desc, float(np.mean(per)) * 1e3
    # fall-back: C port of the two local kernels on a row sample
    from distributed_sddmm_b200 import lib
    from oracle import hnh_oracle as orc
    L = lib()
    N, R = 1 << args.logM, args.R
    mrows = min(N, 1 << 17)
    cap = mrows * args.nnz_per_row
    r = np.empty(cap, np.uint64); c = np.empty(cap, np.uint64); v = np.empty(cap, np.float64)
    n = L.hnh_er_generate_host(args.logM, args.nnz_per_row, SEED, 0, mrows, r.ctypes.data, c.ctypes.data, v.ctypes.data, cap)
    csr = orc.coo_to_csr(mrows, N, r[:n], c[:n], v[:n])
    A = np.full((mrows, R), 0.001); B = np.full((N, R), 0.001)
    vals = np.zeros(n); out = np.zeros((mrows, R))
    ts = []
    for i in range(warmup + steps):
        vals[:] = 0; out[:] = 0
        t0 = time.perf_counter()
        orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, vals, A, B, out)
        ts.append(time.perf_counter() - t0)
    per = np.array(ts[warmup:])
    gf = 4.0 * n * R / per / 1e9
    desc = (f"C port of the reference kernels (oracle/hnh_oracle.c; oracle/_ref not built), first {mrows} of {N} rows "
            f"against the full B, {orc.num_threads()} threads")
    return float(np.mean(gf)), orc.num_threads(), "port", desc, float(np.mean(per)) * 1e3


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    steps, warmup = args.steps, min(args.warmup, 1)
    # keep the whole run within a few minutes whatever K is: cap the number of timed calls
    steps_run = min(steps, 5)
    g, cores, kind, desc, ms = cpu_reference_fusedmm(args, warmup, steps_run, budget_s=90.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": g, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, "cpu", steps_timed=steps_run),
        "cpu_baseline": {"value": g, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": g, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ GPU arm ---------------
def run_native(args):
    import torch
    from distributed_sddmm_b200 import driver as D
    from distributed_sddmm_b200 import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    L = lib()
    rank, world = D.world_init()
    import torch.distributed as dist

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync_barrier():
        L.hnhd_device_synchronize()
        if world > 1:
            L.hnhd_barrier()

    N, R, c = 1 << args.logM, args.R, args.c
    S = D.SpmatLocal.load_er(args.logM, args.nnz_per_row, SEED)
    nnz = S.info()["dist_nnz"]
    alg = D.Algorithm(args.alg, S, R, c)
    info = alg.info()
    A, B = alg.like_A_matrix(0.001), alg.like_B_matrix(0.001)
    Sv, res = alg.like_S_values(1.0), alg.like_S_values(0.0)
    steps_ring = world // c

    def step():
        alg.fusedSpMM(A, B, Sv, res, "A")

    for _ in range(args.warmup):
        step()
    sync_barrier()
    alg.reset_timers()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.hnh_launch_count()
    sync_barrier()
    D.timer_start()
    for _ in range(args.steps):
        step()
    ms_total = D.timer_stop()
    sync_barrier()
    launches = L.hnh_launch_count() - launches0
    clocks = sampler.stop()
    ms = max_over_ranks(ms_total) / args.steps
    flops = 4.0 * nnz * R
    gflops = flops / (ms * 1e-3) / 1e9
    perf = alg.perf()  # collective: averages over ranks

    # ---- roofline of the local kernels (dominant: the fused / SpMM kernel) ----
    comp_ms = perf["Computation Time"] * 1e3 / args.steps
    nnz_rank = float(np.mean(info["nnz_procs"]))
    rows_stationary = alg.dims.localArows * c
    bytes_rank = fusedmm_bytes_per_rank(args.alg, nnz_rank, rows_stationary, steps_ring, R)
    peak, peak_kind = measured_peaks()
    achieved = bytes_rank / (comp_ms * 1e-3) / 1e9 if comp_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_kind": f"of {peak_kind}",
                "kernel": ("fused_row_kernel" if args.alg == "15d_fusion2" else "sddmm_row_kernel + spmm_row_kernel") +
                          f"<{R}>, {steps_ring} launch(es) per step per GPU",
                "kernel_ms_per_step": comp_ms, "algorithmic_bytes_per_step_per_gpu": bytes_rank}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof) and world == 1:
        try:
            roofline["traffic"] = json.load(open(prof)).get(f"{'fused' if args.alg == '15d_fusion2' else 'spmm'}_{R}")
        except Exception:  # noqa: BLE001
            pass
    shift_ms = perf.get("Cyclic Shift Time", 0.0) * 1e3 / args.steps
    repl_ms = perf.get("Replication Time", 0.0) * 1e3 / args.steps

    # ---- the other fusion strategy, for the record (not the headline) ----
    other = None
    if world == 1 and not args.no_other:
        oname = "15d_fusion1" if args.alg == "15d_fusion2" else "15d_fusion2"
        oalg = D.Algorithm(oname, S, R, c)
        oSv, ores = oalg.like_S_values(1.0), oalg.like_S_values(0.0)
        for _ in range(2):
            oalg.fusedSpMM(A, B, oSv, ores, "A")
        sync_barrier()
        D.timer_start()
        for _ in range(max(3, args.steps // 2)):
            oalg.fusedSpMM(A, B, oSv, ores, "A")
        oms = D.timer_stop() / max(3, args.steps // 2)
        other = {oname: {"ms_per_step": oms, "gflops": flops / (oms * 1e-3) / 1e9}}
        del oalg, oSv, ores

    # ---- e2e: per-rank pinned HOST buffers in, result out, copies inside the timed region ----
    shapeA, shapeB = A.shape, B.shape
    hA = torch.full(shapeA, 0.001, dtype=torch.float64).pin_memory()
    hB = torch.full(shapeB, 0.001, dtype=torch.float64).pin_memory()
    hO = torch.empty(shapeA, dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        if args.e2e_pipeline:
            alg.fusedSpMM_host(A, B, Sv, res, hA, hB, hO, "A")
            return
        D.check(L.hnhd_dense_from_host(A.h, hA.data_ptr()), "from_host")
        D.check(L.hnhd_dense_from_host(B.h, hB.data_ptr()), "from_host")
        alg.fusedSpMM(A, B, Sv, res, "A")
        D.check(L.hnhd_dense_to_host(A.h, hO.data_ptr()), "to_host")

    e2e_step()
    sync_barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
    h2d = (shapeA[0] * shapeA[1] + shapeB[0] * shapeB[1]) * 8 * world
    d2h = shapeA[0] * shapeA[1] * 8 * world
    e2e = {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "api": ("per rank: Distributed_Sparse::fusedSpMM_host (pinned host A, B in; result out)" if args.e2e_pipeline else
                   "per rank: DenseMatrix::copy_from_host(A), (B) from pinned memory; fusedSpMM; copy_to_host(A)")}
    del hA, hB, hO

    # ---- CPU baseline on the host cores (rank 0, N = 1 only; bounded sample) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        g, cores, kind, desc, _ = cpu_reference_fusedmm(args, 1, 3)
        cpu = {"value": g, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": desc}

    nvlink = None
    if world > 1:
        try:  # explanatory only: never let it cost the line
            nb = nvlink_bytes_per_rank(args.alg, world, c, alg.dims.localBrows, R)
            nvlink = {"bytes_in_per_gpu_per_step": nb, "peak_gbs": NVLINK_GBS_NOMINAL, "peak_kind": "nominal",
                      "bound_ms": nb / (NVLINK_GBS_NOMINAL * 1e9) * 1e3, "achieved_gbs": nb / (ms * 1e-3) / 1e9}
        except Exception:  # noqa: BLE001
            nvlink = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, "cuda", nnz=nnz, p=world, transport=("nccl" if world > 1 else "self"),
                                      ring_steps=steps_ring, local_rows=alg.dims.localArows),
            "hbm_gbs_achieved_per_gpu": bytes_rank / (ms * 1e-3) / 1e9,
            "roofline": roofline,
            "phase_ms_per_step": {"computation": comp_ms, "cyclic_shift": shift_ms, "replication": repl_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if other:
            line["other"] = other
        if nvlink:
            line["nvlink"] = nvlink
        print(json.dumps(line))
    del alg, S
    D.world_finalize()
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
    ap.add_argument("--c", type=int, default=0, help="replication factor (0 = the tuned default for this GPU count)")
    ap.add_argument("--alg", default="15d_fusion2", choices=["15d_fusion1", "15d_fusion2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true")
    ap.add_argument("--e2e-pipeline", action="store_true",
                    help="e2e leg through fusedSpMM_host (upload / kernel / download pipelined on one rank); "
                         "off until the path has been validated on a GPU")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.c <= 0:
        args.c = default_c(args.alg, world)
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_native(args)


def default_c(alg: str, world: int) -> int:
    """Replication factor per GPU count (measured sweep, profiles/r01_scaling.md)."""
    table = {1: 1, 2: 1, 4: 1, 8: 1}
    return table.get(world, 1)


if __name__ == "__main__":
    sys.exit(main())
