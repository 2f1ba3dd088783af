# Round-2 call D (1 GPU): the device-side setup path -- parity (layouts bit-exact vs oracle/_ref, all operations) with the
# path forced on, then setup time host vs device at BASELINE config 2 and config 3 sizes, then e2e chunk sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multirank_gpu.py -q -rfEs -m gpu -k "device_side_setup" > gpurun_out/r2d_pytest_devsetup.log 2>&1
echo "rc=$?" >> gpurun_out/r2d_pytest_devsetup.log; tail -n 25 gpurun_out/r2d_pytest_devsetup.log
free -g | head -2
timeout 1500 python scripts/setup_bench.py small cfg2 cfg3 > gpurun_out/r2d_setup_bench.jsonl 2> gpurun_out/r2d_setup_bench.err
cat gpurun_out/r2d_setup_bench.jsonl | cut -c1-1200; tail -n 5 gpurun_out/r2d_setup_bench.err
