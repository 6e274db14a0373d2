#!/bin/bash
# bench.py under a few launch-mode / tuning variants; prints value and roofline.frac per variant
mkdir -p gpurun_out
for v in "" "--no-pdl" "--graph" "--graph --no-pdl"; do
  timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-e2e --cpu-budget 0.2 $v 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant [$v] value=%.1f GB/s  ms/step=%.4f  roofline.frac=%.3f  per_qtype=%s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], {k:round(x['frac'],3) for k,x in d['roofline']['per_qtype'].items()}))" || tail -5 gpurun_out/sweep.err
done
