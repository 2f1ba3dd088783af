mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "per_warp" > gpurun_out/pytest_warp.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_warp.log; tail -n 4 gpurun_out/pytest_warp.log
