# Round-2 call B (1 GPU): the two fixed tests + the new empty-block test, the parity leg of bench.py (small, then the
# full default line), then the ncu evidence of the shipped kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -rfEs -m gpu -k "host_operands_multirank or rectangular or empty_block" > gpurun_out/r2b_pytest_fixed.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_pytest_fixed.log; tail -n 15 gpurun_out/r2b_pytest_fixed.log
timeout 300 python bench.py --steps 3 --warmup 3 --logM 14 --nnz-per-row 8 --R 128 > gpurun_out/r2b_bench_small.json 2> gpurun_out/r2b_bench_small.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2b_bench_small.json").read().strip().splitlines()[-1])
    print("small parity:", json.dumps(j.get("parity_check"))[:900])
except Exception as e:
    print("small failed:", e); print(open("gpurun_out/r2b_bench_small.err").read()[-1500:])
PY
/usr/bin/time -v -o gpurun_out/r2b_bench_full.time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err
tail -c 2500 gpurun_out/r2b_bench_full.json; tail -n 5 gpurun_out/r2b_bench_full.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2b_bench_full.time
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other --parity off > gpurun_out/r2b_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_row -s 2 -c 1 -o gpurun_out/r2b_prof_fused128 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other --parity off > gpurun_out/r2b_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
