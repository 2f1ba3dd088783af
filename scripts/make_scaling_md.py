#!/usr/bin/env python
"""profiles/r02_scaling.md from the 8-GPU call's records (gpurun_out/r2g8_*): one table per BASELINE config."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def recs(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return []
    out = []
    for ln in open(p):
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                out.append(json.loads(ln))
            except Exception:  # noqa: BLE001
                pass
    return out


def row(r, note=""):
    if "error" in r:
        return f"| {r.get('alg')} | {r.get('c')} | error: {r['error'][:80]} | | | | |"
    ph = r.get("phase_ms", {})
    phs = ", ".join(f"{k.replace(' Time', '')} {v:.2f}" for k, v in ph.items() if v > 0.004)
    pc = r.get("parity_check") or {}
    par = (f"{pc.get('max_rel_err', float('nan')):.1e} on {pc.get('rows', 0)} rows ({'pass' if pc.get('pass') else 'FAIL'})"
           if pc else "—")
    return (f"| {r['alg']} | {r['c']} | **{r['ms_per_fusedmm']:.2f}** | {r['gflops']:.0f} | {phs} | {par} | {note} |")


out = ["# Round 2 — FusedMM on 8 × B200: BASELINE configs 2–5 (scripts/r2_gpu_8.sh)", "",
       "CUDA-event time of `fusedSpMM(A, B, S, result, Amat)` per call, max over ranks, 10 calls after 3 warm-up; phases are",
       "the library's CUDA-event counters averaged over ranks (they overlap, they do not add up); FLOPs = 4·nnz·R.  `parity`:",
       "one more FusedMM on position-dependent operands on the data plane just timed, row samples of every rank against the C",
       "port of the reference kernels (`bench.py::sample_parity`; tolerance 1e-5).  Rings: copy engines into CUDA-IPC peer",
       "slots (`hnh::PeerRing`) unless noted; set-up on the device.  Raw records: `r02_sweep8_*.jsonl`, `r02_bench_n8.json`.", ""]
hdr = ["| algorithm | c | ms / FusedMM | GFLOP/s | phases (ms) | parity (max rel err) | note |", "|---|---|---|---|---|---|---|"]
for title, files in (("Config 2: ER N=2^20, 32 nnz/row, r=128 (1.5D dense shift)", ["r2g8_sweep_cfg2.jsonl"]),
                     ("Config 5: ER N=2^21, 32 nnz/row, r=128 (1.5D dense shift, local kernel fusion)", ["r2g8_sweep_cfg5.jsonl"]),
                     ("Config 4: ER N=2^20, 32 nnz/row, r=256 (2.5D Cannon dense, c = 2)", ["r2g8_sweep_cfg4.jsonl"]),
                     ("Config 3: ER N=2^22, 64 nnz/row, r=32 (1.5D sparse shift)", ["r2g8_sweep_cfg3.jsonl"])):
    rs = [r for f in files for r in recs(f)]
    out += [f"## {title}", ""] + hdr + [row(r) for r in rs] + [""]
pieces = recs("r2g8_sweep_pieces2.jsonl")
if pieces:
    out += ["## Push pieces (config 2, 15d_fusion2, c = 1): one ring push as n concurrent copies", ""] + hdr + \
        [row(r, "HNH_RING_PIECES=2") for r in pieces] + [""]
apps = recs("r2g8_cfg5_apps.jsonl")
if apps:
    out += ["## Config 5 inside its caller: `benchmark_algorithm(app)` (reference `benchmark_dist.cpp:117-141`)", "",
            "| app | algorithm | ms / call | GFLOP/s (reference FLOP model) | application communication s | perf counters (s, all trials) |", "|---|---|---|---|---|---|"]
    for a in apps:
        out.append(f"| {a.get('app')} | {a.get('alg')} | {a.get('ms_per_call') and round(a['ms_per_call'], 2)} | "
                   f"{a.get('overall_throughput') and round(a['overall_throughput'])} | {a.get('application_communication_time')} | "
                   f"{json.dumps(a.get('perf_stats'))} |")
    out.append("")
b8 = os.path.join(G, "r2g8_bench8.json")
if os.path.exists(b8):
    try:
        j = json.loads(open(b8).read().strip().splitlines()[-1])
        out += ["## `bench.py --gpus 8` (the driver's line)", "",
                f"* value {j['value']:.0f} GFLOP/s, {j['ms_per_step']:.3f} ms per step; phases {json.dumps(j['phase_ms_per_step'])}",
                f"* roofline (kernel time only): {j['roofline']['achieved']:.0f} GB/s per GPU = {j['roofline']['frac']:.2f} of measured peak on "
                f"{j['roofline']['algorithmic_bytes_per_step_per_gpu'] / 1e9:.2f} GB of per-block algorithmic bytes",
                f"* e2e (host operands, pipelined): {j['e2e']['ms_per_step']:.2f} ms per step = {j['e2e']['value']:.0f} GFLOP/s",
                f"* other: {json.dumps(j.get('other'))}", f"* nvlink: {json.dumps(j.get('nvlink'))}",
                f"* run: {json.dumps(j.get('run'))}", f"* parity_check: {json.dumps(j.get('parity_check'))[:600]}", ""]
    except Exception as e:  # noqa: BLE001
        out += [f"(bench8 line unreadable: {e})", ""]
out += ["## Reading", "",
        "* **Correctness at scale.**  Every row above carries a passing `parity_check` (≤ 8.5e-16) computed on the data plane",
        "  that was timed -- copy-engine rings at 8 ranks, NCCL all-gather / reduce-scatter, the 2.5D skew and both of its rings,",
        "  the one-kernel sparse-shift FusedMM; `tests/test_multirank_gpu.py::test_all_operations_match_reference[8]` (all five",
        "  algorithms against `oracle/_ref`, rank by rank) passed on the same box (`r02_pytest_8gpu_rings.log`).",
        "* **Against round 1** (`r01_scaling.md`): config 3 4.70 -> 3.12 ms (one scaled fused kernel instead of an SDDMM and an SpMM",
        "  pass when the CSR block never moves, c = p); config 4 8.39 -> 5.05 ms (both Cannon rings on copy engines instead of NCCL);",
        "  configs 2 and 5 unchanged at 1.58 / 2.98 ms; host-operand e2e at 8 GPUs 9.9 -> 8.3 ms (pipelined); set-up of every record",
        "  on the device (config 3: 0.5 s instead of ~45 s).",
        "* **Why config 2 stops at 1.58 ms on 8 GPUs (0.43 of linear).**  With c = 1 every GPU must take in 7 shards of 128 MiB",
        "  per FusedMM: 0.94 GB at the 690 GB/s one copy-engine push sustains = 1.36 ms (`Cyclic Shift` above), against 0.69 ms for",
        "  linear scaling of the one-GPU step.  The kernels keep pace with the arrivals (1.42 ms for 8 per-block launches; each re-reads",
        "  the row-side factor and reads + writes the accumulator: 7.46 GB per GPU where one launch over all nonzeros would move",
        "  4.63 GB -- `roofline.frac` 0.80 on the former, `frac_on_minimal_bytes` 0.50 on the latter), so the step is paced by NVLink",
        "  ingress, not by the kernels; splitting a push into two concurrent copies does not raise the rate (1.65 ms).  c = 2 would cut",
        "  the ingress to 5 shards (0.67 GB), but its all-gather and reduce-scatter are NCCL collectives serialised against the kernels",
        "  (`Replication` 0.68 ms): 1.98 ms.  Hiding them (copy-engine pushes by row segment, or a fused epilogue that stores partial",
        "  rows into the peer's accumulator) and a multi-block kernel that keeps a row's accumulator in registers across ring slots is",
        "  the open work; the bounds above put its ceiling at about 1.1-1.2 ms (0.6 of linear), not at linear.",
        "* **NCCL send/recv as the ring** (`HNH_RING=nccl`, the literal replacement of the reference's `MPI_Sendrecv`): 6.83 ms per",
        "  FusedMM at 8 GPUs against 1.58 ms with the copy-engine ring (`bench.py`'s `other.nccl_send_recv_ring`).",
        "* **Config 5 in its caller**: one ALS round (`run_cg(1)`: 2 x (1 RHS SpMM + 11 FusedMM + the batched-CG algebra), all on the",
        "  device) takes 92 ms at N = 2^21 on 8 GPUs; 22 FusedMM x 2.98 ms = 66 ms of it.", ""]
open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out))
print("\n".join(out))
