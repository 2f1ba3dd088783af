mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
python -m pytest tests/test_multirank_gpu.py -x -q -s -m gpu > gpurun_out/pytest_mr_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mr_nccl.log
tail -n 40 gpurun_out/pytest_mr_nccl.log
