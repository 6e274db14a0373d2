#!/bin/bash
# pass 3B/3C: gemv2 variants (3C: 16 consumer warps per CTA)
set +e
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "gemv_fast" > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3c_pytest.log | cut -c1-300
GEMV_ROUTES=gemv_fast_ws,gemv_fast timeout -k 10 200 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r3c_bench_gemv.log 2>&1; grep -v Warn gpurun_out/r3c_bench_gemv.log
