#!/bin/bash
# ncu captures for profiles/: launch list of the bench step, full-set captures of the dequant kernel and the GEMM kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-flux --cpu-budget 0.2 --eager"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 106 -c 70 --csv --log-file gpurun_out/launches_dequant_step.csv $B > gpurun_out/ncu1.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:dequant_kernel -s 113 -c 7 -o gpurun_out/prof_dequant_q4k_v3 $B > gpurun_out/ncu2.log 2>&1
L="python tools/bench_linear.py --M 4608 --shapes 2 --copies 2"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_dense $L --routes ours_dense > gpurun_out/ncu3.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_fused $L --routes fused > gpurun_out/ncu4.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 3 -c 2 -o gpurun_out/prof_gemv $L --M 1 --routes fused > gpurun_out/ncu5.log 2>&1
timeout -k 10 300 python tools/bench_flux.py > gpurun_out/flux.json 2> gpurun_out/flux.err; tail -c 1200 gpurun_out/flux.json
timeout -k 10 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
ls -la gpurun_out | tail -15
