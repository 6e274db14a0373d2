#!/bin/bash
# Evidence run after the split-K / LoRA work: full GPU test suite, smoke, bench (N=1, graph + eager + reference arm),
# per-shape Linear benches (M = 4608 eager, M = 512 / 64 kernel time under graph replay), model-shape benches,
# ncu captures of the split-K fused kernel and of the mma.sync GEMV.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|error" | tail -3
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; grep '^{' gpurun_out/bench_final.json | tail -1 | cut -c1-330
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --eager --no-flux --no-e2e --cpu-budget 0.3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager value', d['value'], 'frac', d['roofline']['frac'])"
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 | cut -c1-260
timeout -k 10 300 python tools/bench_linear.py --M 4608 --routes fused dq_mma auto ours_dense cublas > gpurun_out/bench_linear_final.log 2>&1; tail -35 gpurun_out/bench_linear_final.log
timeout -k 10 200 python tools/bench_linear.py --graph --M 512 64 --routes fused dq_mma auto cublas 2>&1 | grep TFLOP > gpurun_out/bench_linear_graph_m512_m64.log; tail -56 gpurun_out/bench_linear_graph_m512_m64.log
timeout -k 10 400 python tools/bench_models.py > gpurun_out/bench_models.log 2> gpurun_out/bench_models.err; cat gpurun_out/bench_models.log | cut -c1-500
L="python tools/bench_linear.py --M 512 --shapes 3 --copies 2"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 6 -c 2 -o gpurun_out/prof_gemm2_fused_splitk $L --routes fused > gpurun_out/ncu_splitk.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemv_mma_kernel -s 6 -c 2 -o gpurun_out/prof_gemv_mma python tools/bench_gemv.py Q4_K > gpurun_out/ncu_gemv.log 2>&1
ls gpurun_out | tail -30
