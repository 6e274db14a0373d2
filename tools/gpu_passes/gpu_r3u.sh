#!/bin/bash
# pass 3U: the LoRA tests on the final binary (mapping helpers moved to produce.cuh)
timeout -k 3 40 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "lora" 2>&1 | tail -2 | cut -c1-120
