#!/bin/bash
# pass 3D: ncu --set full of gemv2 (Q4_K [18432,3072], M = 1) for the stall breakdown + source counters
set +e
mkdir -p gpurun_out
GEMV_ROUTES=gemv_fast timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 10 -c 2 -o gpurun_out/r3d_gemv2_m1 python tools/bench_gemv.py Q4_K > gpurun_out/r3d_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r3d_ncu.log
ncu -i gpurun_out/r3d_gemv2_m1.ncu-rep --page details > gpurun_out/r3d_gemv2_m1_details.txt 2>&1
ncu -i gpurun_out/r3d_gemv2_m1.ncu-rep --page raw --csv > gpurun_out/r3d_gemv2_m1_raw.csv 2>&1
ncu -i gpurun_out/r3d_gemv2_m1.ncu-rep --page source --csv > gpurun_out/r3d_gemv2_m1_source.csv 2>&1
ls -la gpurun_out/r3d*
