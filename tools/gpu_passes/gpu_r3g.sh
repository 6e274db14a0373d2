#!/bin/bash
# pass 3G: Flux-shape step with a rank-32 LoRA on every quantised Linear (in-kernel vs side GEMMs vs unpatched); BASELINE config 4 refresh
set +e
mkdir -p gpurun_out
timeout -k 10 600 python tools/bench_flux.py --steps 6 --ref-steps 0 --lora 32 > gpurun_out/r3g_flux_lora32.json 2> gpurun_out/r3g_flux_lora32.err; echo "flux lora rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r3g_flux_lora32.json') if l.startswith('{')][0]
print('unpatched', d['ms_per_step'], d['lora'])"
timeout -k 10 900 python tools/bench_models.py > gpurun_out/r3g_bench_models.json 2> gpurun_out/r3g_bench_models.err; echo "models rc=$?"; cut -c1-1500 gpurun_out/r3g_bench_models.json
