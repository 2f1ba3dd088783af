mkdir -p gpurun_out
python scripts/e2e_diag.py > gpurun_out/e2e_diag.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/scale_sweep.py > gpurun_out/sweep_n2.log 2>&1; echo "rc=$?" >> gpurun_out/sweep_n2.log
cat gpurun_out/e2e_diag.log; tail -n 4 gpurun_out/bench_n2.log; grep -v "^W\|^\[" gpurun_out/sweep_n2.log | tail -n 12
