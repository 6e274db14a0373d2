#!/bin/bash
# Last evidence run of the round: full GPU suite, smoke, bench (graph / eager / reference arm), model-shape benches,
# host-cost profile, compute-sanitizer memcheck over the newest kernels' tests.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|error" | tail -3
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; grep '^{' gpurun_out/bench_final.json | tail -1 | cut -c1-330
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --eager --no-flux --no-e2e --cpu-budget 0.3 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('eager value', d['value'], 'frac', d['roofline']['frac'])"
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 | cut -c1-260
timeout -k 10 400 python tools/bench_models.py > gpurun_out/bench_models.log 2> gpurun_out/bench_models.err; cat gpurun_out/bench_models.log | cut -c1-420
timeout -k 10 100 python tools/profile_host.py > gpurun_out/profile_host.log 2>&1; head -2 gpurun_out/profile_host.log
timeout -k 10 100 python tools/bench_gemv.py > gpurun_out/bench_gemv.log 2>&1; grep mma gpurun_out/bench_gemv.log
timeout -k 10 240 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_linear.py -m gpu -q -x -k "split_k or lora or small_m or gemv or unaligned" > gpurun_out/memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck.log | tail -3
ls gpurun_out | tail -12
