#!/bin/bash
# pass R: K1 one-tile-per-CTA variants (256 / 128 threads, single-thread mbarrier wait) vs the persistent ring
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
for m in 1 2; do
  GGUFB200_TEST_DEQUANT_MODE=$m timeout -k 10 600 python -m pytest tests/test_gpu_dequant.py -m gpu -q -x > gpurun_out/r2r_pytest_mode$m.log 2>&1; echo "pytest(mode $m) rc=$?"; tail -2 gpurun_out/r2r_pytest_mode$m.log | cut -c1-200
done
for rep in a b; do
for m in 0 1 2; do
  timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-flux --no-e2e --cpu-budget 0.3 --dequant-mode $m > gpurun_out/r2r_bench_mode${m}$rep.json 2> gpurun_out/r2r_bench_mode${m}$rep.err
  python -c "import json; d=json.load(open('gpurun_out/r2r_bench_mode${m}$rep.json')); print('mode $m', round(d['value']), round(d['roofline']['frac'],4), {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, round(d['roofline']['isolated_launch']['frac'],3))"
done
done
