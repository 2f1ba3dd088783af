mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ALGS=15d_fusion1,15d_fusion2 CS=1 $TR --master-port 29531 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n2_ce_cfg2.jsonl
HNH_RING_ALL=0 ALGS=15d_fusion1 CS=1 $TR --master-port 29532 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n2_nccl_cfg2.jsonl
ALGS=15d_sparse LOGM=21 NPR=64 R=32 CS=1 STEPS=5 $TR --master-port 29533 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n2_ce_cfg3.jsonl
HNH_RING_ALL=0 ALGS=15d_sparse LOGM=21 NPR=64 R=32 CS=1 STEPS=5 $TR --master-port 29534 scripts/scale_sweep.py 2>&1 | grep "^{" > gpurun_out/sweep_n2_nccl_cfg3.jsonl
for f in gpurun_out/sweep_n2_ce_cfg2.jsonl gpurun_out/sweep_n2_nccl_cfg2.jsonl gpurun_out/sweep_n2_ce_cfg3.jsonl gpurun_out/sweep_n2_nccl_cfg3.jsonl; do echo "== $f"; cut -c1-330 $f; done
