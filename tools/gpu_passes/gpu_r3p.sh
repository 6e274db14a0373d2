#!/bin/bash
# pass 3P: default LoRA route = side GEMMs: 6 unsynchronised runs at depth 2, the LoRA tests, then the full-depth Flux-shape step with a rank-32 LoRA on all Linears
set +e
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout -k 5 40 python tools/bench_flux.py --depth 2 --depth-single 2 --steps 3 --ref-steps 0 --lora 32 > gpurun_out/r3p_$i.json 2> gpurun_out/r3p_$i.err; echo "side-route run $i rc=$?"
done
timeout -k 10 300 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k lora > gpurun_out/r3p_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3p_pytest.log | cut -c1-200
timeout -k 10 200 python tools/bench_flux.py --steps 6 --ref-steps 0 --lora 32 > gpurun_out/r3p_flux_lora32.json 2> gpurun_out/r3p_flux_lora32.err; echo "full-depth lora rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r3p_flux_lora32.json') if l.startswith('{')][0]
print('unpatched', d['ms_per_step'], d['lora'])"
