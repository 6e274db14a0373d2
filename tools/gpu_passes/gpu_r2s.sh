#!/bin/bash
# pass S: K1 one-tile-per-CTA with 128 vs 64 threads
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
GGUFB200_TEST_DEQUANT_MODE=3 timeout -k 10 600 python -m pytest tests/test_gpu_dequant.py -m gpu -q -x > gpurun_out/r2s_pytest_mode3.log 2>&1; echo "pytest(mode 3) rc=$?"; tail -2 gpurun_out/r2s_pytest_mode3.log | cut -c1-200
for m in 2 3 2 3; do
  timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-flux --no-e2e --cpu-budget 0.3 --dequant-mode $m > gpurun_out/r2s_bench_mode${m}.json 2> gpurun_out/r2s_bench_mode${m}.err
  python -c "import json; d=json.load(open('gpurun_out/r2s_bench_mode${m}.json')); print('mode $m', round(d['value']), round(d['roofline']['frac'],4), {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, round(d['roofline']['isolated_launch']['frac'],3))"
done
K1_MODE=2 timeout -k 10 300 python tools/probe_k1_sizes.py > gpurun_out/r2s_probe_k1_sizes_mode2.log 2>&1; grep -v Warn gpurun_out/r2s_probe_k1_sizes_mode2.log
