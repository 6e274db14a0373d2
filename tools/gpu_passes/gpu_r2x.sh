#!/bin/bash
# pass X: new flag / stream-order tests, K1 over every format and output dtype, full bench with the whole-workload e2e, reference arm on the box
set +e
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_dequant.py tests/test_gpu_linear.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2x_pytest.log | cut -c1-300
timeout -k 10 400 python tools/bench_k1_types.py > gpurun_out/r2x_bench_k1_types.log 2>&1; grep -v Warn gpurun_out/r2x_bench_k1_types.log
( time timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/r2x_bench_reference.json')); print('reference arm', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['config']['tensors_per_step'])"
( time timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/r2x_bench.json')); print(round(d['value']), round(d['roofline']['frac'],4), 'e2e', d['e2e'], 'flux', round(d['flux_step']['ms_per_step'],2), d['cpu_baseline']['value'])"
tail -3 gpurun_out/r2x_bench.err
