#!/bin/bash
# pass U: full evidence run of the round-2 K1 (one tile per CTA, swizzled tensor-map store): tests, smoke, bench, ncu launch list + full capture
set +e
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log | cut -c1-300
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -1 gpurun_out/r2u_smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 --sweep-detail > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python -c "import json; d=json.load(open('gpurun_out/r2u_bench.json')); print(round(d['value']), round(d['roofline']['frac'],4), {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, 'iso', round(d['roofline']['isolated_launch']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'flux', round(d['flux_step']['ms_per_step'],2), d['clocks'])"
timeout -k 10 300 python tools/probe_k1_sizes.py > gpurun_out/r2u_probe_k1_sizes.log 2>&1; grep -v Warn gpurun_out/r2u_probe_k1_sizes.log
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_dequant_step.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-flux --cpu-budget 0.3 > gpurun_out/r2u_ncu_bench1.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dequant_kernel<ggufb200::Block<12>' -s 8 -c 7 -o gpurun_out/r2u_dequant_q4k python bench.py --steps 2 --warmup 3 --no-e2e --no-flux --cpu-budget 0.3 > gpurun_out/r2u_ncu_bench2.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2u_ncu_bench2.log
python tools/ncu_summary.py gpurun_out/r2u_dequant_q4k.ncu-rep > gpurun_out/r2u_dequant_q4k_ncu.txt 2>&1; head -24 gpurun_out/r2u_dequant_q4k_ncu.txt
ls -la gpurun_out/*.ncu-rep
