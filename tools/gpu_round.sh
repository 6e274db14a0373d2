#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + one full capture of the dequant kernel.
mkdir -p gpurun_out
[ "$1" = "probe" ] || timeout -k 10 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -k 10 300 python bench.py --steps 20 --warmup 3 --sweep-detail > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -36 gpurun_out/bench.err | head -40
if [ "$1" = "ncu" ]; then
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 106 -c 70 --csv --log-file gpurun_out/launches_dequant.csv python bench.py --steps 2 --warmup 3 --no-e2e --cpu-budget 0.5 > gpurun_out/ncu_bench1.log 2>&1
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:dequant_kernel -s 113 -c 7 -o gpurun_out/prof_dequant_q4k python bench.py --steps 2 --warmup 3 --no-e2e --cpu-budget 0.5 > gpurun_out/ncu_bench2.log 2>&1
  ls -la gpurun_out/
fi
if [ "$1" = "linear" ] || [ "$2" = "linear" ]; then
  timeout -k 10 600 python tools/bench_linear.py --M 4608 > gpurun_out/bench_linear.log 2>&1; cat gpurun_out/bench_linear.log
fi
if [ "$1" = "probe" ]; then
  timeout -k 10 300 python tools/probe_bw.py 2>&1 | tee gpurun_out/probe_bw.log
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:dequant_kernel -s 113 -c 7 -o gpurun_out/prof_dequant_q4k_v2 python bench.py --steps 2 --warmup 3 --no-e2e --cpu-budget 0.3 > gpurun_out/ncu_bench2.log 2>&1
fi
