# Round-2 first GPU call (1 GPU, ~15-20 min of box time):   scripts/g.sh 1800 scripts/r2_gpu_1.sh
# 1. the paths written at the end of round 1 without GPU time (gated tests), 2. the full GPU suite,
# 3. the bench line with and without the pipelined host-operand e2e leg, 4. per-warp TMA vs the defaults.
mkdir -p gpurun_out
HNH_UNVALIDATED=1 timeout 900 python -m pytest tests -q -rfE -m gpu -k "als_cg_matches or null_and_empty or rectangular or dispatch_table or host_operands or hostpipe or gat or device_ or file_driver or plugged_in or self_check or harness" > gpurun_out/r2_unvalidated.log 2>&1
echo "rc=$?" >> gpurun_out/r2_unvalidated.log; tail -n 40 gpurun_out/r2_unvalidated.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -n 3 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err; tail -c 600 gpurun_out/r2_bench_plain.json
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-pipeline --no-cpu-baseline --no-other > gpurun_out/r2_bench_pipe.json 2> gpurun_out/r2_bench_pipe.err
python - <<'PY'
import json
for n in ("plain", "pipe"):
    try:
        j = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 1), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 1), "e2e ms", round(j["e2e"]["ms_per_step"], 2))
    except Exception as e:
        print(n, "failed:", e)
PY
# HNH_FLAG_TMA_WARP = 128, HNH_FLAG_FORCE_DIRECT = 2, default = 0 (TMA staging where round 1 found it faster)
for f in 0 2 128; do
  HNH_SWEEP_FLAGS=$f HNH_SWEEP_OUT=r2_kernel_sweep_flags$f.json timeout 300 python scripts/kernel_sweep.py r128 r256 > gpurun_out/r2_kernel_sweep_flags$f.log 2>&1
  echo "flags=$f"; cat gpurun_out/r2_kernel_sweep_flags$f.log | tail -n 30
done
