mkdir -p gpurun_out
python scripts/kernel_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/sweep.log
ncu --set full --clock-control none --import-source on -k regex:fused_row -s 3 -c 1 -o gpurun_out/prof_fused128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sddmm_row -s 1 -c 1 -o gpurun_out/prof_sddmm128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --alg 15d_fusion1 >> gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_row -s 1 -c 1 -o gpurun_out/prof_spmm128 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --alg 15d_fusion1 >> gpurun_out/ncu_full.log 2>&1
tail -n 80 gpurun_out/sweep.log
