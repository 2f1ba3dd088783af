# Round-2 eight-GPU call (charged 8x):   gpurun --gpus 8 --timeout 1200 -- "bash scripts/r2_gpu_8.sh"
# Ordered by value: ring parity at 8 ranks, the bench line at N = 8 (sample parity; the full leg is exercised at N = 2
# and by the driver's own scaling run), push-piece sweep, replication sweep, BASELINE configs 5, 4, 3 -- each record
# with a parity_check on the data plane it timed.
mkdir -p gpurun_out
PORT=29817
T() { PORT=$((PORT + 1)); timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT "${@:2}"; }
nvidia-smi topo -m > gpurun_out/r2g8_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multirank_gpu.py -x -q -rfEs -m gpu -k "all_operations and 8" > gpurun_out/r2g8_pytest_mr8.log 2>&1; echo "rc=$?" >> gpurun_out/r2g8_pytest_mr8.log; tail -n 4 gpurun_out/r2g8_pytest_mr8.log
T 600 bench.py --gpus 8 --steps 10 --warmup 3 --parity sample > gpurun_out/r2g8_bench8.json 2> gpurun_out/r2g8_bench8.err; tail -c 1800 gpurun_out/r2g8_bench8.json; tail -n 3 gpurun_out/r2g8_bench8.err
for pieces in 2 4; do
  HNH_RING_PIECES=$pieces ALGS=15d_fusion2 CS=1 PARITY=0 T 300 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_pieces$pieces.jsonl 2> gpurun_out/r2g8_sweep_pieces$pieces.err
  cut -c1-330 gpurun_out/r2g8_sweep_pieces$pieces.jsonl
done
ALGS=15d_fusion2 CS=1,2,4,8 T 400 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg2.jsonl 2> gpurun_out/r2g8_sweep_cfg2.err; cut -c1-420 gpurun_out/r2g8_sweep_cfg2.jsonl
ALGS=15d_fusion1 CS=4,8 T 300 scripts/scale_sweep.py >> gpurun_out/r2g8_sweep_cfg2.jsonl 2>> gpurun_out/r2g8_sweep_cfg2.err; tail -n 2 gpurun_out/r2g8_sweep_cfg2.jsonl | cut -c1-420
# BASELINE config 5: FusedMM stand-alone (with parity) and inside one ALS-CG round, N = 2^21, r = 128
LOGM=21 ALGS=15d_fusion2 CS=1,2 T 300 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg5.jsonl 2> gpurun_out/r2g8_sweep_cfg5.err; cut -c1-420 gpurun_out/r2g8_sweep_cfg5.jsonl
LOGM=21 NPR=32 R=128 APPS=vanilla,als T 300 scripts/app_bench.py > gpurun_out/r2g8_cfg5_apps.jsonl 2> gpurun_out/r2g8_cfg5_apps.err; cut -c1-500 gpurun_out/r2g8_cfg5_apps.jsonl
# BASELINE config 4 (2.5D dense, r = 256, c = 2) and config 3 (1.5D sparse shift, N = 2^22, 64/row, r = 32)
LOGM=20 NPR=32 R=256 ALGS=25d_dense_replicate CS=2 T 300 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg4.jsonl 2> gpurun_out/r2g8_sweep_cfg4.err; cut -c1-500 gpurun_out/r2g8_sweep_cfg4.jsonl
LOGM=22 NPR=64 R=32 ALGS=15d_sparse CS=8,4,1 T 500 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg3.jsonl 2> gpurun_out/r2g8_sweep_cfg3.err; cut -c1-500 gpurun_out/r2g8_sweep_cfg3.jsonl
tail -n 3 gpurun_out/r2g8_*.err | tail -n 30
