# Round-2 8-GPU call (charged 8x; ~10-12 min of box time):   scripts/g.sh --gpus 8 900 scripts/r2_gpu_8.sh
# Ordered by value: parity of the copy-engine rings at 8 ranks (round 1 validated them at 2 and 4), the multi-piece push,
# the bench line with and without the pipelined host-operand leg, BASELINE configs 5, 3, 4.
mkdir -p gpurun_out
PORT=29517
T() { PORT=$((PORT + 1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT "$@"; }
timeout 900 python -m pytest tests/test_multirank_gpu.py -x -q -m gpu -k "all_operations" > gpurun_out/r2_pytest_mr8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_mr8.log; tail -n 3 gpurun_out/r2_pytest_mr8.log
HNH_UNVALIDATED=1 timeout 600 python -m pytest tests/test_dropin.py -q -rfE -m gpu -k "multi_process" > gpurun_out/r2_pytest_cpp_mp.log 2>&1; tail -n 5 gpurun_out/r2_pytest_cpp_mp.log
for pieces in 1 2 4; do
  HNH_RING_PIECES=$pieces ALGS=15d_fusion2,15d_fusion1 CS=1 T scripts/scale_sweep.py > gpurun_out/r2_sweep8_pieces$pieces.log 2>&1
  grep '^{' gpurun_out/r2_sweep8_pieces$pieces.log | cut -c1-400
done
T bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; tail -c 800 gpurun_out/r2_bench8.json
T bench.py --gpus 8 --steps 10 --warmup 3 --e2e-pipeline > gpurun_out/r2_bench8_pipe.json 2> gpurun_out/r2_bench8_pipe.err; tail -c 500 gpurun_out/r2_bench8_pipe.json
# BASELINE config 5: FusedMM stand-alone and inside one ALS-CG round, N = 2^21, r = 128
LOGM=21 NPR=32 R=128 APPS=vanilla,als T scripts/app_bench.py > gpurun_out/r2_cfg5_apps.log 2>&1; grep '^{' gpurun_out/r2_cfg5_apps.log | cut -c1-600
# BASELINE config 4 (2.5D dense, r = 256, c = 2) and config 3 (1.5D sparse shift, N = 2^22, r = 32; the slowest to set up)
LOGM=20 NPR=32 R=256 ALGS=25d_dense_replicate CS=2 T scripts/scale_sweep.py > gpurun_out/r2_sweep8_cfg4.log 2>&1; grep '^{' gpurun_out/r2_sweep8_cfg4.log | cut -c1-400
LOGM=22 NPR=64 R=32 ALGS=15d_sparse CS=1,8 T scripts/scale_sweep.py > gpurun_out/r2_sweep8_cfg3.log 2>&1; grep '^{' gpurun_out/r2_sweep8_cfg3.log | cut -c1-400
