# HEAD check on one GPU: the whole suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -rfEs -m gpu > gpurun_out/r2_final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_pytest.log; tail -n 8 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 700 gpurun_out/r2_final_bench_n1.json
