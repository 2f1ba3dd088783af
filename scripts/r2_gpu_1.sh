# Round-2 one-GPU evidence run:   gpurun --timeout 3000 -- "bash scripts/r2_gpu_1.sh"
# (The round ran these steps in several shorter calls; this is their union, in order.)
mkdir -p gpurun_out
# 1. the whole GPU suite, nothing gated
timeout 1500 python -m pytest tests -q -rfEs -m gpu > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_all.log; tail -n 6 gpurun_out/r2_pytest_all.log
# 2. the driver's bench line (parity leg at full size, pipelined e2e, fusion-1 "other" leg) and the multi-rank control
#    flow of bench.py on two ranks sharing the GPU (testing aid, not a measurement)
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29922 \
    bench.py --gpus 2 --steps 3 --warmup 3 --share-gpu --parity full > gpurun_out/r2_share_full.json 2> gpurun_out/r2_share_full.err; echo "share rc=$?"
# 3. ncu: launch list of the bench command, full capture of the shipped headline kernel, the same kernel at the 8-GPU
#    block shape, DRAM traffic of the narrow-factor kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other --parity off > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_row -s 2 -c 1 -o gpurun_out/r2_prof_fused128 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other --parity off > gpurun_out/r2_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_row -c 2 -o gpurun_out/r2_prof_fused128_block8 \
    python scripts/kernel_sweep.py "cfg2@8" > gpurun_out/r2_ncu_block8.log 2>&1
HNH_SWEEP_OUT=r2_unused.json ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:"split_kernel" --csv --log-file gpurun_out/r2_ncu_small_r.csv \
    python scripts/kernel_sweep.py "cfg3 c=1 r4" "cfg3 c=2 r8" "cfg3 c=4 r16" "cfg3 c=8 r32" > gpurun_out/r2_ncu_small_r.log 2>&1
# 4. per-kernel roofline sweep (default kernels; then direct-load and per-warp TMA at r = 128 / 256)
HNH_SWEEP_OUT=r2_kernel_sweep.json timeout 900 python scripts/kernel_sweep.py > gpurun_out/r2_kernel_sweep.log 2>&1; tail -n 12 gpurun_out/r2_kernel_sweep.log
for f in 2 128; do HNH_SWEEP_FLAGS=$f HNH_SWEEP_OUT=r2_kernel_sweep_flags$f.json timeout 300 python scripts/kernel_sweep.py r128 r256 > gpurun_out/r2_kernel_sweep_flags$f.log 2>&1; done
# 5. set-up time, host path vs device-resident path; one-GPU FusedMM of the other algorithms at a config-3-like size
timeout 1500 python scripts/setup_bench.py small cfg2 cfg3 > gpurun_out/r2_setup_bench.jsonl 2> gpurun_out/r2_setup_bench.err; cut -c1-300 gpurun_out/r2_setup_bench.jsonl
LOGM=20 NPR=64 R=32 ALGS=15d_sparse,15d_fusion2,15d_fusion1 CS=1 timeout 600 python scripts/scale_sweep.py > gpurun_out/r2_sweep1_r32.jsonl 2> gpurun_out/r2_sweep1_r32.err
