"""Setup time, host path vs device path: tuple generation, both redistributions, block splitting and
COO -> CSR of one algorithm object, on ONE rank (one GPU), in a fresh process per path.

    python scripts/setup_bench.py [cfg2 cfg3 ...]      # prints one JSON line per (config, path)

cfg2 = ER N=2^20, 32/row, 15d_fusion2 r=128; cfg3 = ER N=2^22, 64/row (268 M tuples), 15d_sparse r=32."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {"cfg2": (20, 32, "15d_fusion2", 128), "cfg3": (22, 64, "15d_sparse", 32), "small": (16, 16, "15d_fusion2", 32)}


def child(name):
    sys.path.insert(0, ROOT)
    import torch
    torch.cuda.set_device(0)
    from distributed_sddmm_b200 import driver as D
    from distributed_sddmm_b200 import lib
    logM, npr, alg_name, R = CONFIGS[name]
    D.world_init("self")
    lib().hnhd_device_synchronize()
    t0 = time.perf_counter()
    S = D.SpmatLocal.load_er(logM, npr, 0xC0FFEE + 2)
    t1 = time.perf_counter()
    alg = D.Algorithm(alg_name, S, R, 1)
    lib().hnhd_device_synchronize()
    t2 = time.perf_counter()
    # first use moves host-built blocks to HBM: part of the host path's cost
    A, B = alg.like_A_matrix(0.001), alg.like_B_matrix(0.001)
    Sv, res = alg.like_S_values(1.0), alg.like_S_values(0.0)
    alg.fusedSpMM(A, B, Sv, res, "A")
    lib().hnhd_device_synchronize()
    t3 = time.perf_counter()
    print(json.dumps({"config": name, "path": "device" if os.environ.get("HNH_DEVICE_SETUP") == "1" else "host",
                      "nnz": S.info()["dist_nnz"], "generate_s": t1 - t0, "algorithm_ctor_s": t2 - t1,
                      "first_fusedmm_s": t3 - t2, "total_s": t3 - t0, "phases": D.setup_times()}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    for name in (sys.argv[1:] or ["cfg2", "cfg3"]):
        for path in ("0", "1"):
            env = dict(os.environ, HNH_DEVICE_SETUP=path)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=env, capture_output=True, text=True,
                               timeout=3000)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            print(lines[-1] if lines else json.dumps({"config": name, "path": path, "error": (p.stderr or p.stdout)[-800:]}))
            sys.stdout.flush()
