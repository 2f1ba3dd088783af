mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tma_row -s 2 -c 1 -o gpurun_out/prof_tma_sddmm128 python bench.py --alg 15d_fusion1 --steps 1 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/ncu_tma.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
