mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
HNH_SWEEP_OUT=kernel_sweep_direct.json python scripts/kernel_sweep.py "cfg2@1" "r256" "block r128" 2>&1 | grep -E "sddmm |sddmm_b0|fused" > gpurun_out/sweep_direct.log
HNH_SWEEP_FLAGS=64 HNH_SWEEP_OUT=kernel_sweep_tma.json python scripts/kernel_sweep.py "cfg2@1" "r256" "block r128" 2>&1 | grep -E "sddmm |sddmm_b0|fused" > gpurun_out/sweep_tma.log
echo "--- direct"; cat gpurun_out/sweep_direct.log; echo "--- tma"; cat gpurun_out/sweep_tma.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_n1.log; cut -c1-250 gpurun_out/bench_n1.log | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench_v2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/bench_ncu.log 2>&1
