#!/bin/bash
set +e
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_dequant.py -m gpu -q -x -k "src_stable or back_to_back or consecutive" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log | cut -c1-300
