#!/bin/bash
# pass 3F: full verification of the final tree: GPU suite, smoke, bench (default flags, as the driver runs it)
set +e
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3f_pytest.log | cut -c1-300
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f_smoke.log 2>&1; tail -1 gpurun_out/r3f_smoke.log
( time timeout -k 10 900 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err ) 2>&1 | grep real
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r3f_bench.json') if l.startswith('{')][0]
print(round(d['value']), round(d['roofline']['frac'],4), {k: round(v['frac'],3) for k,v in d['roofline']['per_qtype'].items()}, 'e2e', round(d['e2e']['value'],1), 'flux', round(d['flux_step']['ms_per_step'],2), d['clocks'])"
