# Round-2 call E (1 GPU): full GPU suite (own DMMA GEMM, fused sparse-shift FusedMM, device setup at p=1), then the
# sparse-shift FusedMM at a cfg3-like single-GPU size (fused path vs the reference's two passes), e2e chunk sweep,
# and an ncu capture of the fused kernel at the 8-GPU block shape.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -rfEs -m gpu > gpurun_out/r2e_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_pytest_all.log; tail -n 12 gpurun_out/r2e_pytest_all.log
LOGM=20 NPR=64 R=32 ALGS=15d_sparse,15d_fusion2,15d_fusion1 CS=1 timeout 600 python scripts/scale_sweep.py > gpurun_out/r2e_sweep1_r32.jsonl 2> gpurun_out/r2e_sweep1_r32.err
cut -c1-700 gpurun_out/r2e_sweep1_r32.jsonl; tail -n 3 gpurun_out/r2e_sweep1_r32.err
ncu --set full --clock-control none --import-source on -k regex:fused_row -c 2 -o gpurun_out/r2e_prof_fused128_block8 \
    python scripts/kernel_sweep.py "cfg2@8" > gpurun_out/r2e_ncu_block8.log 2>&1
ls -la gpurun_out/r2e_prof_fused128_block8.ncu-rep
