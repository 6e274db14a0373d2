#!/bin/bash
# pass N: plain epilogue restored to the measured-fastest loop, fused tail in its own loop with batched residual loads
set +e
mkdir -p gpurun_out
for t in test_gpu_gemm test_gpu_linear; do
  timeout -k 10 400 python -m pytest tests/$t.py -q -m gpu -x > gpurun_out/r2n_$t.log 2>&1; echo "$t rc=$?"; tail -4 gpurun_out/r2n_$t.log | cut -c1-300
done
echo "== bench_linear bf16 M=4608"; timeout -k 10 300 python tools/bench_linear.py --M 4608 --routes tmem_exact tmem384_exact cublas > gpurun_out/r2n_bench_linear_bf16.log 2>&1; cat gpurun_out/r2n_bench_linear_bf16.log
for mode in "--no-fuse" "--fuse-tail-only" ""; do
  echo "== flux $mode"; timeout -k 10 300 python tools/bench_flux.py --steps 5 --ref-steps 0 $mode > gpurun_out/r2n_flux_${mode//-/}.json 2>> gpurun_out/r2n_flux.err; cut -c1-900 gpurun_out/r2n_flux_${mode//-/}.json | tr ',' '\n' | grep -E "ms_per_step|linear_ms|other_ms"
done
tail -3 gpurun_out/r2n_flux.err
for s in "4096 9216 3072" "4608 3072 15360" "512 3072 12288"; do
  timeout -k 5 60 python tools/probe_lora.py $s 2>&1 | grep -v Warn
done
timeout -k 10 200 python tools/flux_breakdown.py none > gpurun_out/r2n_flux_breakdown_none.txt 2>&1; head -45 gpurun_out/r2n_flux_breakdown_none.txt | cut -c1-200
