#!/bin/bash
# pass Q: ncu --set full of both K1 variants (Q4_K / Q6_K / Q4_0 / Q8_0 at [21504,3072])
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:dequant -o gpurun_out/r2q_k1_modes python tools/ncu_k1_modes.py > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2q_ncu.log
python tools/ncu_summary.py gpurun_out/r2q_k1_modes.ncu-rep > gpurun_out/r2q_k1_modes_ncu.txt 2>&1; grep -c "###" gpurun_out/r2q_k1_modes_ncu.txt
ls -la gpurun_out/*.ncu-rep
