#!/bin/bash
# pass J: gemv2 v3 (producer warp + stage ring): tests, CTAs-per-SM A/B, ncu at M = 1 and M = 8
set +e
mkdir -p gpurun_out
echo "== gemv2 tests"; timeout -k 10 900 python -m pytest tests/test_gpu_linear.py -q -m gpu -k "gemv_fast" > gpurun_out/r2j_gemv2.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2j_gemv2.log | cut -c1-300
echo "== bench_gemv auto"; timeout -k 10 400 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r2j_gemv_bench.log 2>&1; grep -E "gemv_fast|tmem " gpurun_out/r2j_gemv_bench.log
for c in 1 2 3; do
  echo "== bench_gemv ctas=$c"; GGUFB200_ALLOW_TUNING=1 GEMV2_CTAS=$c GEMV_ROUTES=gemv_fast timeout -k 10 300 python tools/bench_gemv.py Q4_K > gpurun_out/r2j_gemv_bench_c$c.log 2>&1; cat gpurun_out/r2j_gemv_bench_c$c.log
done
for m in 1 8; do
echo "== ncu gemv2 M=$m"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 3 -c 1 -f -o /tmp/gv3_$m python tools/bench_linear.py --M $m --shapes 4 --routes gemv_fast > gpurun_out/r2j_ncu_$m.log 2>&1; tail -1 gpurun_out/r2j_ncu_$m.log
ncu -i /tmp/gv3_$m.ncu-rep --page details > gpurun_out/r02_gemv2_v3_m${m}_details.txt 2>/dev/null
ncu -i /tmp/gv3_$m.ncu-rep --page source --csv > gpurun_out/r02_gemv2_v3_m${m}_source.csv 2>/dev/null
done
