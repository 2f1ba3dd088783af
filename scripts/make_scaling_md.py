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
open(os.path.join(ROOT, "profiles", "r02_scaling.md"), "w").write("\n".join(out))
print("\n".join(out))
