"""Drop-in checks of the C++ surface: the reference's own driver source compiles UNCHANGED against
include/hnh/compat + libhnh_b200.so, and the stand-alone drivers run on the GPU."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "distributed_sddmm_b200")


def test_reference_driver_compiles_unchanged():
    if not os.path.exists("/root/reference/bench_erdos_renyi.cpp"):
        pytest.skip("/root/reference is not on this box")
    from distributed_sddmm_b200 import build
    exe = build.build_reference_driver()
    assert exe and os.path.exists(exe)
    # bench_file.cpp, bench_heatmap.cpp, scratch.cpp; bench_erdos_renyi.cpp + the reference's own benchmark_dist.cpp
    for other in ("bench_file_reference_main", "bench_heatmap_reference_main", "scratch_reference_main",
                  "bench_er_reference_harness"):
        assert os.path.exists(os.path.join(PKG, other)), other
    import torch
    if not torch.cuda.is_available():
        # no CPU fallback: the program must refuse to run without a GPU
        p = subprocess.run([exe, "6", "4", "15d", "8", "1", "/tmp/hnh_dropin.json"], capture_output=True, text=True)
        assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["bench_er_reference_main", "bench_er"])
def test_cpp_drivers_run_on_gpu(exe, tmp_path):
    path = os.path.join(PKG, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not built")
    out = tmp_path / "records.json"
    p = subprocess.run([path, "12", "8", "15d", "32", "1", str(out)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    text = out.read_text().strip().rstrip(",")
    records = json.loads("[" + text + "]")  # the reference appends `record,` per benchmark
    assert [r["alg_name"] for r in records] == ["15d_fusion1", "15d_fusion2"]
    for r in records:
        assert r["fused"] is True and r["num_trials"] == 5 and r["overall_throughput"] > 0
        assert r["alg_info"]["m"] == 4096 and r["alg_info"]["r"] == 32 and r["alg_info"]["p"] == 1


@pytest.mark.gpu
def test_reference_file_driver_runs_on_gpu(tmp_path):
    """The reference's bench_file.cpp (MatrixMarket input, unfused SDDMM + SpMM on the 1.5D sparse-shift algorithm),
    compiled unchanged."""
    path = os.path.join(PKG, "bench_file_reference_main")
    if not os.path.exists(path):
        pytest.skip("bench_file_reference_main not built")
    import numpy as np
    from oracle import hnh_oracle as orc
    rows, cols, _ = orc.er_tuples(10, 6, 5)
    mtx = tmp_path / "er.mtx"
    with open(mtx, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"1024 1024 {len(rows)}\n")
        np.savetxt(f, np.column_stack([rows + 1, cols + 1, np.ones(len(rows))]), fmt="%d %d %g")
    out = tmp_path / "records.json"
    p = subprocess.run([path, str(mtx), "15d", "32", "1", str(out), "vanilla"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    records = json.loads("[" + out.read_text().strip().rstrip(",") + "]")
    assert [r["alg_name"] for r in records] == ["15d_sparse"] and records[0]["fused"] is False
    assert records[0]["alg_info"]["nnz"] == len(rows) and records[0]["overall_throughput"] > 0


@pytest.mark.gpu
def test_reference_self_check_program_runs_on_gpu(tmp_path):
    """The reference's scratch.cpp -- its only self-check: squared-norm fingerprints of sddmmA / spmmA / spmmB on
    dummyInitialize inputs with the 1.5D sparse-shift algorithm, then a GAT forward pass -- compiled unchanged.
    The reference prints the fingerprints without expected values; here they are compared with
    scipy on the global matrices (X[i, j] = i * R + j, distributed_sparse.h:322-346)."""
    path = os.path.join(PKG, "scratch_reference_main")
    if not os.path.exists(path):
        pytest.skip("scratch_reference_main not built")
    import numpy as np
    import scipy.sparse as sp
    from oracle import hnh_oracle as orc
    logM, R = 8, 8
    N = 1 << logM
    rows, cols, _ = orc.er_tuples(logM, 6, 5)
    mtx = tmp_path / "er.mtx"
    with open(mtx, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{N} {N} {len(rows)}\n")
        np.savetxt(f, np.column_stack([rows + 1, cols + 1, np.ones(len(rows))]), fmt="%d %d %g")
    p = subprocess.run([path, str(mtx), str(R), "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    got = {}
    for line in p.stdout.splitlines():
        if "Fingerprint:" in line:
            k, v = line.split("Fingerprint:")
            got[k.strip()] = float(v)
    X = orc.dummy_matrix(0, N, R)
    S = sp.csr_matrix((np.ones(len(rows)), (rows.astype(np.int64), cols.astype(np.int64))), shape=(N, N))
    S.sort_indices()
    ri = np.repeat(np.arange(N), np.diff(S.indptr))
    want = {"SDDMM": float((np.einsum("ij,ij->i", X[ri], X[S.indices]) ** 2).sum()),
            "SpMMA": float(((S @ X) ** 2).sum()), "SpMMB": float(((S.T @ X) ** 2).sum())}
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-5 * v, (k, got[k], v)  # the program prints 6 significant digits


@pytest.mark.gpu
@pytest.mark.parametrize("alg,R", [("15d_fusion2", 16), ("15d_fusion1", 16), ("15d_sparse", 8)])
def test_reference_als_source_runs_unchanged_and_matches_the_device_als(alg, R):
    """BASELINE.json north_star: "drops into ... als_conjugate_gradients unchanged".  als_reference_main is this repo's
    small driver + the REFERENCE's als_conjugate_gradients.cpp compiled unchanged (Eigen expressions, the host loop
    over data() in scale_matrix_rows, raw MPI_Allreduce: host-access mode).  One alternating round of its cg_optimizer
    on coordinate-determined inputs must give the numbers of the library's device-resident ALS on the same inputs."""
    import numpy as np
    import torch
    from distributed_sddmm_b200 import driver as D
    from oracle import ref
    exe = os.path.join(PKG, "als_reference_main")
    if not os.path.exists(exe):
        pytest.skip("als_reference_main is built only where /root/reference exists")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    logM, npr = 9, 6
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    p = subprocess.run([exe, str(logM), str(npr), alg, str(R), "1", "10"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    rec = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["host_access_mode"] is True
    # the same run through the library's own ALS (csrc/als_conjugate_gradients.cpp: all algebra in CUDA kernels)
    D.world_init("self")
    try:
        N = 1 << logM
        S = D.SpmatLocal.load_er(logM, npr)
        a = D.Algorithm(alg, S, R, 1)
        mats = [ref.pattern(N, R, salt) / R for salt in (1, 2, 3, 4)]
        (before, after), A, B = D.als_run(a, *mats, 1, 10)
        del a, S
    finally:
        D.world_finalize(destroy_process_group=False)
    assert abs(rec["residual_before"] - before) <= 1e-12 * abs(before)
    assert after < before and abs(rec["residual_after"] - after) <= 1e-9 * abs(after), (rec, before, after)
    assert abs(rec["fingerprint_A"] - float(np.sum(A * A))) <= 1e-9 * rec["fingerprint_A"]
    assert abs(rec["fingerprint_B"] - float(np.sum(B * B))) <= 1e-9 * rec["fingerprint_B"]


@pytest.mark.gpu
def test_reference_benchmark_harness_runs_on_gpu(tmp_path):
    """bench_erdos_renyi.cpp AND the reference's own benchmark_dist.cpp (benchmark_algorithm: algorithm selection,
    benchmark inputs, five-trial loop, FLOP model, JSON record), both compiled unchanged, on this library's classes.
    (`MPI_Barrier` of the compat layer drains the GPU streams, so the harness' wall clock covers the device work.)"""
    path = os.path.join(PKG, "bench_er_reference_harness")
    if not os.path.exists(path):
        pytest.skip("bench_er_reference_harness not built")
    out = tmp_path / "records.json"
    p = subprocess.run([path, "12", "8", "15d", "32", "1", str(out)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    records = json.loads("[" + out.read_text().strip().rstrip(",") + "]")
    assert [r["alg_name"] for r in records] == ["15d_fusion1", "15d_fusion2"]
    for r in records:
        assert r["fused"] is True and r["num_trials"] == 5 and r["overall_throughput"] > 0 and r["elapsed"] > 0
        assert r["alg_info"]["m"] == 4096 and r["alg_info"]["r"] == 32 and r["alg_info"]["p"] == 1
        assert set(r["perf_stats"]) >= {"Computation Time"}


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["bench_er_reference_main", "bench_er_reference_harness"])
def test_cpp_driver_multi_process_nccl(exe, tmp_path):
    """The reference's driver as `mpirun -n 2` would start it: two processes, one GPU each, the world built from the
    torchrun-style environment and the NCCL id exchanged through HNH_NCCL_ID_FILE (hnh_world_init_from_env)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not put two ranks on one device)")
    path = os.path.join(PKG, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not built")
    out = tmp_path / "records.json"
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", HNH_NCCL_ID_FILE=str(tmp_path / "nccl.id"))
        procs.append(subprocess.Popen([path, "12", "8", "15d", "32", "1", str(out)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-1500:] for l in logs)
    records = json.loads("[" + out.read_text().strip().rstrip(",") + "]")  # rank 0 writes
    assert [r["alg_name"] for r in records] == ["15d_fusion1", "15d_fusion2"]
    for r in records:
        assert r["alg_info"]["p"] == 2 and r["num_trials"] == 5 and r["overall_throughput"] > 0
