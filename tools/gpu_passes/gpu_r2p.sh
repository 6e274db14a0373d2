#!/bin/bash
# pass P: K1 one-tile-per-CTA variant (hardware-ordered tiles) vs the persistent ring, parity of the variant
set +e
mkdir -p gpurun_out
export GGUFB200_ALLOW_TUNING=1
GGUFB200_TEST_DEQUANT_MODE=1 timeout -k 10 600 python -m pytest tests/test_gpu_dequant.py -m gpu -q -x > gpurun_out/r2p_pytest_np.log 2>&1; echo "pytest(np) rc=$?"; tail -3 gpurun_out/r2p_pytest_np.log | cut -c1-200
K1_MODE=1 timeout -k 10 300 python tools/probe_k1_sizes.py > gpurun_out/r2p_probe_k1_sizes_np.log 2>&1; grep -v Warn gpurun_out/r2p_probe_k1_sizes_np.log
for m in 0 1; do
  timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-flux --no-e2e --cpu-budget 0.3 --dequant-mode $m > gpurun_out/r2p_bench_mode$m.json 2> gpurun_out/r2p_bench_mode$m.err
  python -c "import json; d=json.load(open('gpurun_out/r2p_bench_mode$m.json')); print('mode $m', d['value'], d['roofline']['frac'], {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, d['roofline'].get('isolated_launch'))"
done
