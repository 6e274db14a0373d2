#!/bin/bash
# pass O: SRC_STABLE / W_STABLE (packed tiles fetched under the previous kernel's tail): parity suite, K1 size sweep, A/B of the bench
set +e
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2o_pytest.log | cut -c1-300
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; tail -2 gpurun_out/r2o_smoke.log
timeout -k 10 300 python tools/probe_k1_sizes.py > gpurun_out/r2o_probe_k1_sizes.log 2>&1; cat gpurun_out/r2o_probe_k1_sizes.log | grep -v Warn
GEMV_ROUTES=gemv_fast_ws,gemv_fast timeout -k 10 200 python tools/bench_gemv.py Q4_K Q5_K > gpurun_out/r2o_bench_gemv.log 2>&1; cat gpurun_out/r2o_bench_gemv.log | grep -v Warn
timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-flux --no-e2e --cpu-budget 0.3 --no-src-stable > gpurun_out/r2o_bench_plain.json 2> gpurun_out/r2o_bench_plain.err
python -c "import json; d=json.load(open('gpurun_out/r2o_bench_plain.json')); print('plain', d['value'], d['roofline']['frac'], {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()})"
timeout -k 10 500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2o_bench.json')); print('stable', d['value'], d['roofline']['frac'], {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, d['e2e']['value'], d['flux_step']['ms_per_step'])"
tail -3 gpurun_out/r2o_bench.err
