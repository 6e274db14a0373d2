#!/bin/bash
# pass W: gemv2 with one row tile per CTA (hardware-ordered) vs the persistent grid
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
for c in 0 100; do
  echo "== GEMV2_CTAS=$c"
  GEMV2_CTAS=$c GEMV_ROUTES=gemv_fast_ws,gemv_fast timeout -k 10 200 python tools/bench_gemv.py Q4_K > gpurun_out/r2w_bench_gemv_$c.log 2>&1; grep -v Warn gpurun_out/r2w_bench_gemv_$c.log
done
