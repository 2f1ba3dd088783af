# Round-2 call G (1 GPU): the per-kernel roofline sweep over the block shapes of BASELINE.json's configs (round-2
# kernels), then DRAM traffic / L2 hit rate of the narrow-factor kernels (r = 4, 8, 16, 32) under ncu metrics.
mkdir -p gpurun_out
HNH_SWEEP_OUT=r2g_kernel_sweep.json timeout 900 python scripts/kernel_sweep.py > gpurun_out/r2g_kernel_sweep.log 2>&1; tail -n 75 gpurun_out/r2g_kernel_sweep.log
HNH_SWEEP_OUT=r2g_unused.json ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:"split_kernel" --csv --log-file gpurun_out/r2g_ncu_small_r.csv \
    python scripts/kernel_sweep.py "cfg3 c=1 r4" "cfg3 c=2 r8" "cfg3 c=4 r16" "cfg3 c=8 r32" > gpurun_out/r2g_ncu_small_r.log 2>&1
wc -l gpurun_out/r2g_ncu_small_r.csv
