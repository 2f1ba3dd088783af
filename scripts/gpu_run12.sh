mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
python -m pytest tests/test_multirank_gpu.py -x -q -s -m gpu -k "8" > gpurun_out/pytest_mr8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mr8.log
tail -n 5 gpurun_out/pytest_mr8.log
$TR --master-port 29521 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n8_cfg2.jsonl; cat gpurun_out/sweep_n8_cfg2.jsonl | cut -c1-330
$TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 2>&1 | grep "^{" > gpurun_out/bench_n8.json; cut -c1-400 gpurun_out/bench_n8.json
HNH_RING=nccl ALGS=15d_fusion2 CS=1 $TR --master-port 29523 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n8_ncclring.jsonl; cut -c1-330 gpurun_out/sweep_n8_ncclring.jsonl
ALGS=25d_dense_replicate R=256 CS=2 $TR --master-port 29524 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n8_cfg4.jsonl; cut -c1-400 gpurun_out/sweep_n8_cfg4.jsonl
ALGS=15d_fusion2 LOGM=21 CS=1 $TR --master-port 29525 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n8_cfg5.jsonl; cut -c1-330 gpurun_out/sweep_n8_cfg5.jsonl
ALGS=15d_sparse LOGM=22 NPR=64 R=32 CS=1,8 STEPS=5 $TR --master-port 29526 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n8_cfg3.jsonl; cut -c1-400 gpurun_out/sweep_n8_cfg3.jsonl
