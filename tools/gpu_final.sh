#!/bin/bash
# Final evidence run of a round: full GPU test suite, smoke, bench (N=1), per-shape Linear bench, ncu recaptures.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|error" | tail -2
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; grep '^{' gpurun_out/bench_final.json | tail -1 | cut -c1-330
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --eager --no-flux --no-e2e --cpu-budget 0.3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager value', d['value'], 'frac', d['roofline']['frac'])"
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 | cut -c1-260
timeout -k 10 400 python tools/bench_linear.py --M 4608 --routes fused dq_mma ours_dense cublas ref_chain > gpurun_out/bench_linear_final.log 2>&1; tail -40 gpurun_out/bench_linear_final.log
timeout -k 10 300 python tools/bench_linear.py --M 512 --routes fused dq_mma auto cublas > gpurun_out/bench_linear_m512.log 2>&1
timeout -k 10 400 python tools/bench_models.py > gpurun_out/bench_models.log 2> gpurun_out/bench_models.err; cat gpurun_out/bench_models.log | cut -c1-500
L="python tools/bench_linear.py --M 4608 --shapes 2 --copies 2"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm3_kernel -s 6 -c 2 -o gpurun_out/prof_gemm3_dense $L --routes ours_dense > gpurun_out/ncu3.log 2>&1
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_fused_final $L --routes fused > gpurun_out/ncu4.log 2>&1
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_flux_step.csv python tools/bench_flux.py --depth 2 --depth-single 2 --steps 1 --ref-steps 0 > gpurun_out/ncu6.log 2>&1
ls gpurun_out | tail -30
