#!/bin/bash
# round-2 GPU pass F: converged-warp MMA issue (gemm4 + gemm3): full GPU suite + timings
set +e
mkdir -p gpurun_out
for f in test_gpu_dequant test_gpu_gemm test_gpu_linear test_gpu_flux; do
  timeout -k 10 1500 python -m pytest tests/$f.py -q -m gpu -x > gpurun_out/r2f_$f.log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/r2f_$f.log | head -3
  grep -E "^FAILED" gpurun_out/r2f_$f.log | head -12
done
R="tmem tmem384 tmem_exact tmem384_exact fused dq_mma ours_dense cublas"
echo "== bench_linear bf16 M=4608"; timeout -k 10 600 python tools/bench_linear.py --graph --M 4608 --routes $R > gpurun_out/r2f_bl_bf16_m4608.log 2>&1; cat gpurun_out/r2f_bl_bf16_m4608.log | cut -c1-150
echo "== bench_linear f16 M=4608"; timeout -k 10 300 python tools/bench_linear.py --graph --act f16 --M 4608 --shapes 2 3 --routes tmem tmem384 dq_mma ours_dense cublas > gpurun_out/r2f_bl_f16_m4608.log 2>&1; cat gpurun_out/r2f_bl_f16_m4608.log | cut -c1-150
echo "== bench_linear bf16 M=512 / 64 / 1024 / 2048"; timeout -k 10 400 python tools/bench_linear.py --graph --M 2048 1024 512 64 --shapes 0 3 6 --routes tmem tmem384 fused dq_mma cublas > gpurun_out/r2f_bl_midm.log 2>&1; cat gpurun_out/r2f_bl_midm.log | cut -c1-150
echo "== bench_gemv"; timeout -k 10 400 python tools/bench_gemv.py Q4_K > gpurun_out/r2f_gemv.log 2>&1; grep -v "tmem_spans\|tmem_exact" gpurun_out/r2f_gemv.log
