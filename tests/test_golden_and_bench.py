"""Golden fixtures stay in sync with the reference build; bench.py's reference arm honours the JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import mp_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_golden_files_match_fresh_reference_run():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built on this box")
    from tests.test_multirank_gpu import CASES
    checked = 0
    for p, c in ((2, CASES[2][0]), (4, CASES[4][5]), (2, CASES[2][7])):
        path = os.path.join(ROOT, "tests", "golden", f"{c['name']}_p{p}.npz")
        assert os.path.exists(path), f"missing golden file {path}: run scripts/make_golden.py"
        fresh, src = U.reference_for(c, p)
        assert src == "oracle/_ref"
        gold = U.load_golden(path, p)
        for f, g in zip(fresh, gold):
            assert np.array_equal(f["S_rows"], g["S_rows"]) and np.array_equal(f["ST_cols"], g["ST_cols"])
            for t in range(len(f["ops"])):
                for k in ("A", "B", "values"):
                    assert np.array_equal(f["ops"][t][k], g["ops"][t][k]), (c["name"], t, k)
        checked += 1
    assert checked == 3


def test_bench_reference_arm_contract():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--logM", "12", "--nnz-per-row", "8", "--R", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one line on stdout"
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "GFLOP/s" and rec["higher_is_better"] is True
    assert rec["value"] > 0 and rec["dtype"] == "f64" and "workload" in rec["config"]
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]


def test_bench_native_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_bench_byte_models():
    """The byte models behind bench.py's `roofline` and `nvlink` objects at BASELINE config 2."""
    import bench
    nnz, m, r = 33_553_937, 1 << 20, 128
    # one GPU: one overwrite-fused launch over the whole matrix = 37.05 GB
    assert abs(bench.algorithmic_bytes("fused", nnz, m, r, beta0=True) / 1e9 - 37.05) < 0.01
    assert bench.fusedmm_bytes_per_rank("15d_fusion2", nnz, m, 1, r) == bench.algorithmic_bytes("fused", nnz, m, r, beta0=True)
    # eight GPUs, c = 1: seven shards of 2^17 x 128 doubles arrive per FusedMM (0.875 GiB)
    shard = 8 * (1 << 17) * 128
    assert bench.nvlink_bytes_per_rank("15d_fusion2", 8, 1, 1 << 17, 128) == 7 * shard
    assert bench.nvlink_bytes_per_rank("15d_fusion1", 8, 1, 1 << 17, 128) == 15 * shard  # output ring goes all the way round
    # c = p: no ring, all-gather in + reduce-scatter in
    assert bench.nvlink_bytes_per_rank("15d_fusion2", 8, 8, 1 << 17, 128) == 14 * shard
    assert bench.nvlink_bytes_per_rank("15d_fusion2", 1, 1, 1 << 20, 128) == 0
