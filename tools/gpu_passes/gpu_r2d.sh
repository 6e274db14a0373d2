#!/bin/bash
# round-2 GPU pass D: why is gemm4 v1 slow?  span-major staging vs 2-D tensor map, and ncu --set full of one GEMM-mode and one GEMV-mode launch.
set +e
mkdir -p gpurun_out
echo "== bench_linear bf16 M=4608 shapes 2,3: canonical vs span-major staging"
timeout -k 10 300 python tools/bench_linear.py --graph --M 4608 --shapes 2 3 --routes tmem tmem_spans tmem384 tmem384_spans > gpurun_out/r2d_bl_spans.log 2>&1; cat gpurun_out/r2d_bl_spans.log | cut -c1-150
echo "== bench_gemv"; timeout -k 10 300 python tools/bench_gemv.py Q4_K > gpurun_out/r2d_gemv.log 2>&1; cat gpurun_out/r2d_gemv.log
echo "== ncu gemm mode"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o gpurun_out/r02_gemm4_v1_m4608 python tools/bench_linear.py --M 4608 --shapes 2 --routes tmem > gpurun_out/r2d_ncu1.log 2>&1; tail -3 gpurun_out/r2d_ncu1.log
echo "== ncu gemv mode"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o gpurun_out/r02_gemm4_v1_gemv python tools/bench_linear.py --M 1 --shapes 4 --routes tmem > gpurun_out/r2d_ncu2.log 2>&1; tail -3 gpurun_out/r2d_ncu2.log
ls -la gpurun_out/*.ncu-rep
