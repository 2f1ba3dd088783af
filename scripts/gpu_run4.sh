mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
HNH_SPLIT_MAX_R=32 python -m pytest tests -x -q -m gpu -k "matches_oracle or beta0" >> gpurun_out/pytest_gpu.log 2>&1; echo "pytest(split32) rc=$?" >> gpurun_out/pytest_gpu.log
HNH_SPLIT_MAX_R=32 HNH_SWEEP_OUT=kernel_sweep_split32.json python scripts/kernel_sweep.py r32 r16 r8 r4 > gpurun_out/sweep_split.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_split.log
