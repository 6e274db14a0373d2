#!/bin/bash
# pass T: K1 output tile through a swizzled tensor-map store (static chunk order) vs the linear tile with rotated chunk order
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
GGUFB200_TEST_TUNING="3=1" timeout -k 10 600 python -m pytest tests/test_gpu_dequant.py tests/test_gpu_linear.py -m gpu -q -x > gpurun_out/r2t_pytest_swz.log 2>&1; echo "pytest(swz) rc=$?"; tail -2 gpurun_out/r2t_pytest_swz.log | cut -c1-200
for v in "3=0" "3=1" "3=0" "3=1"; do
  timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-flux --no-e2e --cpu-budget 0.3 --tuning $v > gpurun_out/r2t_bench_$v.json 2> gpurun_out/r2t_bench_$v.err
  python -c "import json; d=json.load(open('gpurun_out/r2t_bench_$v.json')); print('tuning $v', round(d['value']), round(d['roofline']['frac'],4), {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, round(d['roofline']['isolated_launch']['frac'],3))"
done
