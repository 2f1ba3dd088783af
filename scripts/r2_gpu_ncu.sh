# Round-2 profiling call (1 GPU):   scripts/g.sh 900 scripts/r2_gpu_ncu.sh
# Launch list of the default bench command (share of each kernel in the step) and one full capture of the
# dominant kernel; read here with `ncu -i gpurun_out/<name>.ncu-rep --page raw --csv`, summarise under profiles/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_row -s 2 -c 1 -o gpurun_out/r2_prof_fused128 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/r2_ncu_fused.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"fused_row|sddmm_row|spmm_row|tma_row" -c 12 \
    --csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_traffic.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_*.csv | tail -5
