mkdir -p gpurun_out
python -m pytest tests -x -q -s -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "nproc=|passed|failed|rc=|Error|error" gpurun_out/pytest_gpu.log | tail -n 12
