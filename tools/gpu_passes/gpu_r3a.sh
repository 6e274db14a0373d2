#!/bin/bash
# pass 3A: mid-M (333 / 512 / 1024 tokens) -- the planner's choice vs forced item sizes / no K ranges vs cuBLAS on a dense weight
set +e
mkdir -p gpurun_out
timeout -k 10 900 python tools/bench_linear.py --graph --M 512 1024 --routes tmem tmem192 tmem384 tmem_ns tmem192_ns tmem384_ns dq_mma ours_dense cublas --nk 3072 3072 9216 3072 12288 3072 3072 12288 4096 4096 10240 4096 4096 10240 > gpurun_out/r3a_bench_linear_midm.log 2>&1
grep -v Warn gpurun_out/r3a_bench_linear_midm.log | cut -c1-120
