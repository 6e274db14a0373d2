#!/bin/bash
# pass Y (2 GPUs): torchrun bench at N=2 (replicas, whole-workload e2e), two-device tests, compute-sanitizer memcheck of the K1 tests
set +e
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-flux > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2y_bench_n2.json')); print(round(d['value']), round(d['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['numa'])"
tail -3 gpurun_out/r2y_bench_n2.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2y_bench_ref_n2.json 2> gpurun_out/r2y_bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/r2y_bench_ref_n2.json
timeout -k 10 300 python -m pytest tests/test_gpu_multi_device.py -m gpu -q > gpurun_out/r2y_pytest_multi.log 2>&1; tail -2 gpurun_out/r2y_pytest_multi.log
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dequant.py -m gpu -q -x -k "ragged or src_stable or back_to_back or consecutive or unaligned or empty" > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2y_memcheck.log | cut -c1-200
