# Round-2 two-GPU call (charged 2x):   gpurun --gpus 2 --timeout 1500 -- "bash scripts/r2_gpu_2.sh"
# NCCL + copy-engine-ring transport on real GPUs for the paths the one-GPU box can only run over gloo, and the
# bench line at N = 2 with the full parity leg (reference on rank 0, rows handed out over gloo).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2g2_gpus.txt
timeout 900 python -m pytest tests -q -rfEs -m gpu -k "multi_process_nccl or (all_operations and 2) or (device_side_setup and 2) or (host_operands_multirank and 2) or (als_cg_matches and 2)" > gpurun_out/r2g2_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2g2_pytest.log; tail -n 12 gpurun_out/r2g2_pytest.log
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g2_bench2.json 2> gpurun_out/r2g2_bench2.err
echo "bench wall: $(( $(date +%s) - S )) s"; tail -c 2500 gpurun_out/r2g2_bench2.json; tail -n 5 gpurun_out/r2g2_bench2.err
