mkdir -p gpurun_out
python scripts/kernel_sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?" >> gpurun_out/sweep.log
tail -n 80 gpurun_out/sweep.log
