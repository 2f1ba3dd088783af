mkdir -p gpurun_out
python scripts/contend.py > gpurun_out/contend.log 2>&1; cat gpurun_out/contend.log | tail -8
python scripts/e2e_diag.py 2>&1 | tail -4
