#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
timeout -k 10 300 python tools/bench_linear.py --M 4608 --routes fused dq_mma ours_dense 2>&1 | tee gpurun_out/bench_linear3.log
timeout -k 10 200 python tools/bench_linear.py --M 1 8 --routes auto --shapes 1 4 2>&1 | tee gpurun_out/bench_gemv.log
L="python tools/bench_linear.py --M 4608 --shapes 2 --copies 2"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_dense $L --routes ours_dense > gpurun_out/ncu3.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_fused_staged $L --routes fused > gpurun_out/ncu4.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 3 -c 2 -o gpurun_out/prof_gemv python tools/bench_linear.py --M 1 --shapes 4 --copies 2 --routes auto > gpurun_out/ncu5.log 2>&1
timeout -k 10 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 900 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
