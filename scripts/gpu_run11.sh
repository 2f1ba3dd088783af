mkdir -p gpurun_out
python -m pytest tests/test_multirank_gpu.py -x -q -s -m gpu > gpurun_out/pytest_mr_ipc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mr_ipc.log
tail -n 12 gpurun_out/pytest_mr_ipc.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/scale_sweep.py > gpurun_out/sweep_n2_ipc.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_n2_ipc.log
grep -v "^W\|^\[\|^\*\|OMP_NUM" gpurun_out/sweep_n2_ipc.log | tail -n 8
HNH_RING=nccl ALGS=15d_fusion2 CS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/scale_sweep.py 2>&1 | grep "^{" | tail -2
