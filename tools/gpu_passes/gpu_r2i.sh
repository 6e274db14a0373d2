#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== gemv2 tests"; timeout -k 10 900 python -m pytest tests/test_gpu_linear.py -q -m gpu -k "gemv_fast" > gpurun_out/r2i_gemv2.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2i_gemv2.log | cut -c1-300
echo "== bench_gemv"; timeout -k 10 400 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r2i_gemv_bench.log 2>&1; grep -v "tmem_exact" gpurun_out/r2i_gemv_bench.log
echo "== ncu gemv2"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 3 -c 1 -f -o /tmp/gv2 python tools/bench_linear.py --M 8 --shapes 4 --routes gemv_fast > gpurun_out/r2i_ncu.log 2>&1; tail -1 gpurun_out/r2i_ncu.log
ncu -i /tmp/gv2.ncu-rep --page raw --csv > gpurun_out/r02_gemv2_m8_raw.csv 2>/dev/null
ncu -i /tmp/gv2.ncu-rep --page details > gpurun_out/r02_gemv2_m8_details.txt 2>/dev/null
ncu -i /tmp/gv2.ncu-rep --page source --csv > gpurun_out/r02_gemv2_m8_source.csv 2>/dev/null
echo "== models (from pass H, if absent)"; cut -c1-600 gpurun_out/r2h_models.json 2>/dev/null
