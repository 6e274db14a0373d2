#!/bin/bash
# pass 3H: LoRA diagnostics (each shape in its own process under a short timeout)
set +e
mkdir -p gpurun_out
for s in "4608 3072 3072" "4608 18432 3072" "1 18432 3072" "1 9216 3072" "512 12288 3072" "4096 21504 3072" "4096 3072 15360" "512 3072 12288" "1 3072 256"; do
  timeout -k 5 60 python tools/probe_lora.py $s 2>&1 | grep -v Warn; echo "  rc=$? ($s)"
done 2>&1 | tee gpurun_out/r3h_probe_lora.log
timeout -k 5 150 python tools/bench_flux.py --depth 2 --depth-single 2 --steps 3 --ref-steps 0 --lora 32 > gpurun_out/r3h_flux_small_lora.json 2> gpurun_out/r3h_flux_small_lora.err; echo "small flux lora rc=$?"; cut -c1-600 gpurun_out/r3h_flux_small_lora.json | tr ',' '\n' | grep -i "lora\|in_kernel\|side" | head
