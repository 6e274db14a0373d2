#!/bin/bash
# pass K: stmatrix epilogue of gemm4 + gemv2 with hoisted loads: tests, benches, ncu of the tile-384 kernel, Flux step with a LoRA
set +e
mkdir -p gpurun_out
for t in test_gpu_gemm test_gpu_linear; do
  timeout -k 10 1200 python -m pytest tests/$t.py -q -m gpu -x > gpurun_out/r2k_$t.log 2>&1; echo "$t rc=$?"; tail -4 gpurun_out/r2k_$t.log | cut -c1-300
done
echo "== bench_gemv"; GEMV_ROUTES=gemv_fast timeout -k 10 300 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r2k_gemv_bench.log 2>&1; cat gpurun_out/r2k_gemv_bench.log
echo "== bench_linear bf16 M=4608"; timeout -k 10 600 python tools/bench_linear.py --M 4608 --routes tmem_exact tmem384_exact tmem384 ours_dense cublas > gpurun_out/r2k_bench_linear_bf16.log 2>&1; cat gpurun_out/r2k_bench_linear_bf16.log
echo "== ncu gemm4 tile384 exact"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o /tmp/g4e python tools/bench_linear.py --M 4608 --shapes 2 --routes tmem384_exact > gpurun_out/r2k_ncu_g4.log 2>&1; tail -1 gpurun_out/r2k_ncu_g4.log
ncu -i /tmp/g4e.ncu-rep --page details > gpurun_out/r02_gemm4_v4_tile384_details.txt 2>/dev/null
ncu -i /tmp/g4e.ncu-rep --page source --csv > gpurun_out/r02_gemm4_v4_tile384_source.csv 2>/dev/null
echo "== flux exact + lora 32"; timeout -k 10 900 python tools/bench_flux.py --steps 5 --ref-steps 0 --lora 32 > gpurun_out/r2k_flux_lora.json 2> gpurun_out/r2k_flux_lora.err; cut -c1-1500 gpurun_out/r2k_flux_lora.json; tail -3 gpurun_out/r2k_flux_lora.err
