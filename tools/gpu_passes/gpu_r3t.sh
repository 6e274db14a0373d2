#!/bin/bash
# pass 3T: last check of the final tree: LoRA tests with the in-kernel route as the default again, smoke
set +e
mkdir -p gpurun_out
timeout -k 5 50 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "lora or patch" > gpurun_out/r3t_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3t_pytest.log | cut -c1-160
timeout -k 5 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
