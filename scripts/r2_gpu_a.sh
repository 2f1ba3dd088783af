# Round-2 call A (1 GPU): everything gated in round 1 + the full GPU suite in ONE pytest run (no -x),
# then the bench line with the pipelined host-operand leg, then per-warp TMA vs defaults.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt
HNH_UNVALIDATED=1 timeout 1000 python -m pytest tests -q -rfEs -m gpu > gpurun_out/r2a_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_pytest_all.log; tail -n 60 gpurun_out/r2a_pytest_all.log
timeout 500 python bench.py --steps 10 --warmup 3 --e2e-pipeline --no-cpu-baseline --no-other > gpurun_out/r2a_bench_pipe.json 2> gpurun_out/r2a_bench_pipe.err
tail -c 1500 gpurun_out/r2a_bench_pipe.json; tail -n 5 gpurun_out/r2a_bench_pipe.err
for f in 0 2 128; do
  HNH_SWEEP_FLAGS=$f HNH_SWEEP_OUT=r2a_kernel_sweep_flags$f.json timeout 300 python scripts/kernel_sweep.py r128 r256 > gpurun_out/r2a_kernel_sweep_flags$f.log 2>&1
  echo "flags=$f"; tail -n 14 gpurun_out/r2a_kernel_sweep_flags$f.log
done
