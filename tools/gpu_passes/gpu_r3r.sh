#!/bin/bash
# pass 3R: in-kernel LoRA route after the stage-ownership fix of the gemm4 producers: unsynchronised forwards, 5 runs + the LoRA tests
set +e
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  GGUFB200_LORA_NOSYNC=1 timeout -k 3 22 python tools/bench_flux.py --depth 2 --depth-single 2 --steps 3 --ref-steps 0 --lora 32 --lora-in-kernel > gpurun_out/r3r_$i.json 2> gpurun_out/r3r_$i.err; echo "fixed in-kernel run $i rc=$?"
done
timeout -k 5 60 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k lora > gpurun_out/r3r_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3r_pytest.log | cut -c1-160
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r3r_1.json') if l.startswith('{')]
print(d[0]['lora'] if d else 'no json')"
