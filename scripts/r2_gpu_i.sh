# Round-2 call I (1 GPU): HEAD check -- multi-rank control flow of bench.py on shared-GPU ranks (must now print its line),
# the GPU suite, and the default bench line.
mkdir -p gpurun_out
T2() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
T2 200 29921 bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity full --logM 16 > gpurun_out/r2i_share_small.json 2> gpurun_out/r2i_share_small.err; echo "share small rc=$?"; tail -c 1500 gpurun_out/r2i_share_small.json; grep "bench +" gpurun_out/r2i_share_small.err | tail -n 3
T2 300 29922 bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity full > gpurun_out/r2i_share_full.json 2> gpurun_out/r2i_share_full.err; echo "share full rc=$?"; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2i_share_full.json").read().strip().splitlines()[-1])
    print("parity:", json.dumps(j["parity_check"])[:700]); print("other:", j.get("other")); print("e2e:", j["e2e"]["ms_per_step"], "run:", j["run"])
except Exception as e:
    print("failed", e)
PY
timeout 1200 python -m pytest tests -q -rfEs -m gpu > gpurun_out/r2i_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest_all.log; tail -n 6 gpurun_out/r2i_pytest_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; tail -c 1200 gpurun_out/r2i_bench_n1.json
