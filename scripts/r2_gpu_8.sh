# Round-2 eight-GPU call (charged 8x):   gpurun --gpus 8 --timeout 780 -- "bash scripts/r2_gpu_8.sh"
# Ordered by value: ring parity at 8 ranks, the bench line at N = 8 (sample parity; the full leg is exercised at N = 1,
# on shared-GPU ranks and by the driver's own scaling run), push pieces, replication sweep, BASELINE configs 5, 4, 3 --
# each record with a parity_check on the data plane it timed.  Tight per-command timeouts: a stall must not eat the budget.
mkdir -p gpurun_out
PORT=29817
T() { PORT=$((PORT + 1)); timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT "${@:2}"; }
timeout 240 python -m pytest tests/test_multirank_gpu.py -x -q -rfEs -m gpu -k "all_operations and 8" > gpurun_out/r2g8_pytest_mr8.log 2>&1; echo "rc=$?" >> gpurun_out/r2g8_pytest_mr8.log; tail -n 4 gpurun_out/r2g8_pytest_mr8.log
T 200 bench.py --gpus 8 --steps 10 --warmup 3 --parity sample > gpurun_out/r2g8_bench8.json 2> gpurun_out/r2g8_bench8.err; echo "bench8 rc=$?"; tail -c 2000 gpurun_out/r2g8_bench8.json; grep "bench +" gpurun_out/r2g8_bench8.err | tail -n 3
HNH_RING_PIECES=2 ALGS=15d_fusion2 CS=1 PARITY=0 T 120 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_pieces2.jsonl 2> gpurun_out/r2g8_sweep_pieces2.err; cut -c1-330 gpurun_out/r2g8_sweep_pieces2.jsonl
ALGS=15d_fusion2 CS=1,2,4,8 T 200 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg2.jsonl 2> gpurun_out/r2g8_sweep_cfg2.err; cut -c1-420 gpurun_out/r2g8_sweep_cfg2.jsonl
ALGS=15d_fusion1 CS=8 T 120 scripts/scale_sweep.py >> gpurun_out/r2g8_sweep_cfg2.jsonl 2>> gpurun_out/r2g8_sweep_cfg2.err; tail -n 1 gpurun_out/r2g8_sweep_cfg2.jsonl | cut -c1-420
# BASELINE config 5: FusedMM stand-alone (with parity) and inside one ALS-CG round, N = 2^21, r = 128
LOGM=21 ALGS=15d_fusion2 CS=1 T 150 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg5.jsonl 2> gpurun_out/r2g8_sweep_cfg5.err; cut -c1-420 gpurun_out/r2g8_sweep_cfg5.jsonl
LOGM=21 NPR=32 R=128 APPS=als TRIALS=2 T 150 scripts/app_bench.py > gpurun_out/r2g8_cfg5_apps.jsonl 2> gpurun_out/r2g8_cfg5_apps.err; cut -c1-500 gpurun_out/r2g8_cfg5_apps.jsonl
# BASELINE config 4 (2.5D dense, r = 256, c = 2) and config 3 (1.5D sparse shift, N = 2^22, 64/row, r = 32)
LOGM=20 NPR=32 R=256 ALGS=25d_dense_replicate CS=2 T 150 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg4.jsonl 2> gpurun_out/r2g8_sweep_cfg4.err; cut -c1-500 gpurun_out/r2g8_sweep_cfg4.jsonl
LOGM=22 NPR=64 R=32 ALGS=15d_sparse CS=8,4 T 200 scripts/scale_sweep.py > gpurun_out/r2g8_sweep_cfg3.jsonl 2> gpurun_out/r2g8_sweep_cfg3.err; cut -c1-500 gpurun_out/r2g8_sweep_cfg3.jsonl
for f in gpurun_out/r2g8_*.err; do echo "== $f"; grep -v "^$\|OMP_NUM_THREADS\|\*\*\*\*" $f | tail -n 3; done 2>/dev/null | tail -n 40
