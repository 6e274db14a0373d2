#!/bin/bash
# pass L: per-warp pipelined epilogue + fused element-wise tail; where does the LoRA-in-kernel Flux step stall? (one shape per process, short timeouts)
set +e
mkdir -p gpurun_out
rm -f gpurun_out/r2l_probe_lora.log
for s in "300 512 512" "1 18432 3072" "512 9216 3072" "512 3072 12288" "4096 9216 3072" "4608 21504 3072" "4608 3072 15360"; do
  timeout -k 5 60 python tools/probe_lora.py $s >> gpurun_out/r2l_probe_lora.log 2>&1; echo "shape $s rc=$?" | tee -a gpurun_out/r2l_probe_lora.log
done
grep -v Warning gpurun_out/r2l_probe_lora.log | tail -40
for t in test_gpu_gemm test_gpu_linear; do
  timeout -k 10 400 python -m pytest tests/$t.py -q -m gpu -x > gpurun_out/r2l_$t.log 2>&1; echo "$t rc=$?"; tail -4 gpurun_out/r2l_$t.log | cut -c1-300
done
echo "== bench_linear bf16 M=4608"; timeout -k 10 300 python tools/bench_linear.py --M 4608 --routes tmem_exact tmem384_exact ours_dense cublas > gpurun_out/r2l_bench_linear_bf16.log 2>&1; cat gpurun_out/r2l_bench_linear_bf16.log
echo "== bench_gemv"; GEMV_ROUTES=gemv_fast timeout -k 10 200 python tools/bench_gemv.py Q4_K > gpurun_out/r2l_gemv_bench.log 2>&1; cat gpurun_out/r2l_gemv_bench.log
echo "== flux fused / unfused"; timeout -k 10 300 python tools/bench_flux.py --steps 5 --ref-steps 1 > gpurun_out/r2l_flux_fused.json 2> gpurun_out/r2l_flux.err; cut -c1-1800 gpurun_out/r2l_flux_fused.json; tail -3 gpurun_out/r2l_flux.err
timeout -k 10 300 python tools/bench_flux.py --steps 5 --ref-steps 0 --no-fuse > gpurun_out/r2l_flux_unfused.json 2>> gpurun_out/r2l_flux.err; cut -c1-1800 gpurun_out/r2l_flux_unfused.json
