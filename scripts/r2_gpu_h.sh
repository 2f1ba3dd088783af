# Round-2 call H (1 GPU): the multi-rank control flow of bench.py (two ranks SHARING cuda:0 over the gloo transport --
# a testing aid, not a measurement) with phase markers, to locate the stall seen at N = 2; then call G's kernel sweep.
mkdir -p gpurun_out
T2() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
T2 240 29911 bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity sample --logM 16 > gpurun_out/r2h_share_small.json 2> gpurun_out/r2h_share_small.err; echo "small rc=$?"; grep "bench +" gpurun_out/r2h_share_small.err | tail -n 12; tail -c 600 gpurun_out/r2h_share_small.json
T2 300 29912 bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity sample > gpurun_out/r2h_share_sample.json 2> gpurun_out/r2h_share_sample.err; echo "sample rc=$?"; grep "bench +" gpurun_out/r2h_share_sample.err | tail -n 12; tail -c 600 gpurun_out/r2h_share_sample.json
T2 420 29913 bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity full > gpurun_out/r2h_share_full.json 2> gpurun_out/r2h_share_full.err; echo "full rc=$?"; grep "bench +" gpurun_out/r2h_share_full.err | tail -n 12; tail -c 900 gpurun_out/r2h_share_full.json
bash scripts/r2_gpu_g.sh
