#!/bin/bash
# round-2 GPU pass E: cheap (CTA-scope) remote arrives + TMA-store epilogue: tests, timings, then one ncu capture exported to text on the box
set +e
mkdir -p gpurun_out
for f in test_gpu_gemm test_gpu_linear; do
  timeout -k 10 1500 python -m pytest tests/$f.py -q -m gpu -x > gpurun_out/r2e_$f.log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/r2e_$f.log | head -3
  grep -E "^FAILED" gpurun_out/r2e_$f.log | head -12
done
R="tmem tmem384 tmem_exact tmem384_exact fused dq_mma cublas"
echo "== bench_linear bf16 M=4608"; timeout -k 10 600 python tools/bench_linear.py --graph --M 4608 --routes $R > gpurun_out/r2e_bl_bf16_m4608.log 2>&1; cat gpurun_out/r2e_bl_bf16_m4608.log | cut -c1-150
echo "== bench_linear bf16 M=512 / 64"; timeout -k 10 300 python tools/bench_linear.py --graph --M 512 64 --shapes 0 3 6 --routes tmem tmem_exact fused dq_mma cublas > gpurun_out/r2e_bl_m512.log 2>&1; cat gpurun_out/r2e_bl_m512.log | cut -c1-150
echo "== bench_gemv"; timeout -k 10 400 python tools/bench_gemv.py Q4_K Q8_0 > gpurun_out/r2e_gemv.log 2>&1; grep -v tmem_spans gpurun_out/r2e_gemv.log
echo "== ncu gemm mode (tmem384)"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o /tmp/g4a python tools/bench_linear.py --M 4608 --shapes 2 --routes tmem384 > gpurun_out/r2e_ncu1.log 2>&1; tail -2 gpurun_out/r2e_ncu1.log
ncu -i /tmp/g4a.ncu-rep --page raw --csv > gpurun_out/r02_gemm4_tile384_m4608_raw.csv 2>/dev/null
ncu -i /tmp/g4a.ncu-rep --page details > gpurun_out/r02_gemm4_tile384_m4608_details.txt 2>/dev/null
ncu -i /tmp/g4a.ncu-rep --page source --csv > gpurun_out/r02_gemm4_tile384_m4608_source.csv 2>/dev/null
echo "== ncu gemv mode"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o /tmp/g4b python tools/bench_linear.py --M 1 --shapes 4 --routes tmem > gpurun_out/r2e_ncu2.log 2>&1; tail -2 gpurun_out/r2e_ncu2.log
ncu -i /tmp/g4b.ncu-rep --page raw --csv > gpurun_out/r02_gemm4_gemv_m1_raw.csv 2>/dev/null
ncu -i /tmp/g4b.ncu-rep --page details > gpurun_out/r02_gemm4_gemv_m1_details.txt 2>/dev/null
ncu -i /tmp/g4b.ncu-rep --page source --csv > gpurun_out/r02_gemm4_gemv_m1_source.csv 2>/dev/null
ls -la gpurun_out/ | head -30; du -sh gpurun_out
