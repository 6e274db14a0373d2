#!/bin/bash
# round-2 GPU pass A: probe the new TMEM-fed kernel first (mixed f16 x bf16 UMMA, A from TMEM), then the suites and first timings.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
echo "== probe"; timeout -k 10 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "tmem_fused_all_types and Q4_K" > gpurun_out/r2a_probe.log 2>&1; echo "probe rc=$?"; tail -5 gpurun_out/r2a_probe.log
echo "== tmem tests"; timeout -k 10 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "tmem" > gpurun_out/r2a_tmem.log 2>&1; echo "tmem rc=$?"; tail -15 gpurun_out/r2a_tmem.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
echo "== all gpu tests"; timeout -k 10 1500 python -m pytest tests -q -m gpu > gpurun_out/r2a_all.log 2>&1; echo "all rc=$?"; tail -25 gpurun_out/r2a_all.log
echo "== bench_linear M=4608"; timeout -k 10 400 python tools/bench_linear.py --graph --M 4608 --routes tmem tmem384 tmem_generic fused dq_mma cublas > gpurun_out/r2a_bl_m4608.log 2>&1; tail -45 gpurun_out/r2a_bl_m4608.log
echo "== bench_linear M=512"; timeout -k 10 300 python tools/bench_linear.py --graph --M 512 --shapes 0 3 6 --routes tmem fused dq_mma cublas > gpurun_out/r2a_bl_m512.log 2>&1; tail -15 gpurun_out/r2a_bl_m512.log
echo "== bench_gemv"; timeout -k 10 300 python tools/bench_gemv.py Q4_K Q8_0 > gpurun_out/r2a_gemv.log 2>&1; tail -40 gpurun_out/r2a_gemv.log
