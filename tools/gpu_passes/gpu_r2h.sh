#!/bin/bash
# round-2 GPU pass H: K3 v2 (gemv2) tests + timings, models (SD3.5 / T5), remaining suite
set +e
mkdir -p gpurun_out
echo "== gemv2 tests"; timeout -k 10 600 python -m pytest tests/test_gpu_linear.py -q -m gpu -x -k "gemv_fast" > gpurun_out/r2h_gemv2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2h_gemv2.log | cut -c1-300
for f in test_gpu_gemm test_gpu_linear test_gpu_flux; do
  timeout -k 10 1500 python -m pytest tests/$f.py -q -m gpu > gpurun_out/r2h_$f.log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/r2h_$f.log | head -3
  grep -E "^FAILED" gpurun_out/r2h_$f.log | head -12
done
echo "== bench_gemv"; timeout -k 10 400 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r2h_gemv_bench.log 2>&1; cat gpurun_out/r2h_gemv_bench.log
echo "== models"; timeout -k 10 900 python tools/bench_models.py t5 sd35 > gpurun_out/r2h_models.json 2> gpurun_out/r2h_models.err; cut -c1-700 gpurun_out/r2h_models.json; tail -2 gpurun_out/r2h_models.err
echo "== flux exact"; timeout -k 10 600 python tools/bench_flux.py --ref-steps 0 > gpurun_out/r2h_flux_exact.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2h_flux_exact.json'));print({k:d[k] for k in ('ms_per_step','linear_ms','other_ms','numerics')})"
