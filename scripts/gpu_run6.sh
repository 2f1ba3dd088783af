mkdir -p gpurun_out
python -m pytest tests/test_multirank_gpu.py -x -q -s -m gpu > gpurun_out/pytest_mr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mr.log
tail -n 60 gpurun_out/pytest_mr.log
