#!/bin/bash
# round-2 GPU pass C: full GPU suite (one process per file: a device fault would otherwise poison the rest), then first timings.
set +e
mkdir -p gpurun_out
for f in test_gpu_dequant test_gpu_gemm test_gpu_linear test_gpu_flux test_gpu_multi_device; do
  timeout -k 10 1500 python -m pytest tests/$f.py -q -m gpu > gpurun_out/r2c_$f.log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/r2c_$f.log | head -3
  grep -E "^FAILED" gpurun_out/r2c_$f.log | head -12
done
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c_smoke.log
R="tmem tmem384 tmem_exact tmem384_exact tmem_generic fused dq_mma cublas"
echo "== bench_linear bf16 M=4608"; timeout -k 10 600 python tools/bench_linear.py --graph --M 4608 --routes $R > gpurun_out/r2c_bl_bf16_m4608.log 2>&1; cat gpurun_out/r2c_bl_bf16_m4608.log | cut -c1-150
echo "== bench_linear f16 M=4608 (3 shapes)"; timeout -k 10 300 python tools/bench_linear.py --graph --act f16 --M 4608 --shapes 0 2 3 --routes tmem tmem384 tmem_exact dq_mma cublas > gpurun_out/r2c_bl_f16_m4608.log 2>&1; cat gpurun_out/r2c_bl_f16_m4608.log | cut -c1-150
echo "== bench_linear bf16 M=512 / 64"; timeout -k 10 300 python tools/bench_linear.py --graph --M 512 64 --shapes 0 3 6 --routes tmem tmem_exact fused dq_mma cublas > gpurun_out/r2c_bl_m512.log 2>&1; cat gpurun_out/r2c_bl_m512.log | cut -c1-150
echo "== bench_gemv"; timeout -k 10 400 python tools/bench_gemv.py Q4_K Q8_0 Q5_K > gpurun_out/r2c_gemv.log 2>&1; cat gpurun_out/r2c_gemv.log
