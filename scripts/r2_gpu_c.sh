# Round-2 call C (1 GPU): whole GPU suite after the SDDMM-epilogue (Hadamard fold) change, then the default bench line
# (parity leg at full size, pipelined e2e, fusion-1 "other" leg).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -rfEs -m gpu -x > gpurun_out/r2c_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_pytest_all.log; tail -n 12 gpurun_out/r2c_pytest_all.log
S=$(date +%s)
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err
echo "bench wall: $(( $(date +%s) - S )) s"
tail -c 3000 gpurun_out/r2c_bench_full.json; tail -n 5 gpurun_out/r2c_bench_full.err
