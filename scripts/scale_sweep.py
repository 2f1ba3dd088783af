#!/usr/bin/env python
"""FusedMM time and phase breakdown over (algorithm, c) at the launched world size (torchrun).
Prints one JSON record per configuration on rank 0; summarised in profiles/."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import PARITY_RTOL, SEED, fusedmm_bytes_per_rank, sample_parity  # noqa: E402
from distributed_sddmm_b200 import driver as D  # noqa: E402
from distributed_sddmm_b200 import lib  # noqa: E402

logM = int(os.environ.get("LOGM", "20"))
npr = int(os.environ.get("NPR", "32"))
R = int(os.environ.get("R", "128"))
algs = os.environ.get("ALGS", "15d_fusion2,15d_fusion1").split(",")
steps = int(os.environ.get("STEPS", "10"))

torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
L = lib()
rank, world = D.world_init()
import torch.distributed as dist  # noqa: E402


def mx(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


S = D.SpmatLocal.load_er(logM, npr, SEED)
nnz = S.info()["dist_nnz"]
cs = [c for c in (1, 2, 4, 8) if world % c == 0 and c <= world]
if os.environ.get("CS"):
    cs = [int(x) for x in os.environ["CS"].split(",")]
for name in algs:
    for c in cs:
        if name.startswith("25d"):
            s = int(round((world / c) ** 0.5))
            if s * s * c != world:
                continue
        try:
            alg = D.Algorithm(name, S, R, c)
        except RuntimeError as e:
            if rank == 0:
                print(json.dumps({"alg": name, "c": c, "error": str(e)[:200]}), flush=True)
            continue
        A, B = alg.like_A_matrix(0.001), alg.like_B_matrix(0.001)
        Sv, res = alg.like_S_values(1.0), alg.like_S_values(0.0)
        for _ in range(3):
            alg.fusedSpMM(A, B, Sv, res, "A")
        L.hnhd_device_synchronize(); L.hnhd_barrier()
        alg.reset_timers()
        D.timer_start()
        for _ in range(steps):
            alg.fusedSpMM(A, B, Sv, res, "A")
        ms = mx(D.timer_stop()) / steps
        L.hnhd_device_synchronize(); L.hnhd_barrier()
        perf = alg.perf()
        info = alg.info()
        # correctness of the data plane just timed, at this size (row samples of every rank against the C port of the
        # reference kernels; Cannon layouts need the alignment shifts the benchmark loop leaves out)
        parity = None
        if os.environ.get("PARITY", "1") != "0":
            try:
                _, parity = sample_parity(logM, npr, R, alg, A, B, Sv, res, rank, world, shifts=name.startswith("25d"),
                                          rows_per_block=int(os.environ.get("PARITY_ROWS", "1024")))
                parity["pass"] = bool(parity["max_rel_err"] <= PARITY_RTOL)
            except Exception as e:  # noqa: BLE001
                parity = {"error": f"{type(e).__name__}: {e}", "pass": False}
        if rank == 0:
            rec = {"alg": name, "p": world, "c": c, "R": R, "logM": logM, "ms_per_fusedmm": ms,
                   "gflops": 4.0 * nnz * R / ms / 1e6,
                   "phase_ms": {k: v * 1e3 / steps for k, v in perf.items()},
                   "nnz_max_over_ranks": max(info["nnz_procs"]), "nnz_mean": float(np.mean(info["nnz_procs"])),
                   "ring": info.get("ring"), "parity_check": parity, "setup_s": D.setup_times(reset=True)}
            if name in ("15d_fusion1", "15d_fusion2"):
                by = fusedmm_bytes_per_rank(name, rec["nnz_mean"], alg.dims.localArows * c, world // c, R)
                rec["kernel_GBps_per_gpu"] = by / (rec["phase_ms"]["Computation Time"] * 1e-3) / 1e9
            print(json.dumps(rec), flush=True)
        del alg, A, B, Sv, res
D.world_finalize()
