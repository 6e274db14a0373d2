#!/bin/bash
# pass V: ncu --set full of the 7 Q4_K launches of one bench step (K1, one tile per CTA + swizzled tensor-map store)
set +e
mkdir -p gpurun_out
timeout -k 10 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dequant_kernel<ggufb200::Block<.int.12>' -s 8 -c 7 -o gpurun_out/r2u_dequant_q4k python bench.py --steps 2 --warmup 3 --no-e2e --no-flux --cpu-budget 0.3 > gpurun_out/r2u_ncu_bench2.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2u_ncu_bench2.log
python tools/ncu_summary.py gpurun_out/r2u_dequant_q4k.ncu-rep > gpurun_out/r2u_dequant_q4k_ncu.txt 2>&1; head -24 gpurun_out/r2u_dequant_q4k_ncu.txt
ncu -i gpurun_out/r2u_dequant_q4k.ncu-rep --page details > gpurun_out/r2u_dequant_q4k_ncu_details.txt 2>&1
ls -la gpurun_out/*.ncu-rep
