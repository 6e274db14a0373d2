#!/bin/bash
# round-2 GPU pass B: one process per probe (a device fault is sticky), then the suites.
set +e
mkdir -p gpurun_out
P="timeout -k 10 120 python tools/probe_tmem.py"
{
echo "== hardware feature probe (tools/probe_ts.cu): 0 st/ld, 1 TS f16xf16, 2 TS f16 x bf16 (mixed), 3 TS bf16xbf16, 4 SS f16"
for v in 0 4 1 3 2; do timeout -k 5 60 tools/probe_ts $v 2>&1 | tail -1; done
echo "== A f16 M=300 (tt=192)"; $P f16 0 2>&1 | tail -2
echo "== B f16 M=100 (tt=128)"; $P f16 0 100 264 1024 2>&1 | tail -2
echo "== C f16 M=8 (tt=32, split-K)"; $P f16 0 8 512 2048 2>&1 | tail -2
echo "== D bf16 WCAST"; $P bf16 2000 2>&1 | tail -2
echo "== E bf16 mixed f16 x bf16"; $P bf16 0 2>&1 | tail -2
echo "== F f16 tile384 M=400"; $P f16 400 400 264 1024 2>&1 | tail -2
echo "== G f16 generic Q8_0"; $P f16 200 300 264 1024 Q8_0 2>&1 | tail -2
echo "== H f16 big"; $P f16 0 4608 3072 3072 2>&1 | tail -2
} > gpurun_out/r2b_probes.log 2>&1
cat gpurun_out/r2b_probes.log
if grep -q "illegal" gpurun_out/r2b_probes.log; then
  echo "== sanitizer on the first failing probe"
  if grep -A1 "== A" gpurun_out/r2b_probes.log | grep -q illegal; then ARGS="f16 0"; else ARGS="bf16 0"; fi
  timeout -k 10 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/probe_tmem.py $ARGS > gpurun_out/r2b_sanitizer.log 2>&1
  grep -E "Illegal|at 0x|in /|by thread|ERROR SUMMARY|Invalid" gpurun_out/r2b_sanitizer.log | head -20
fi
echo "== tmem tests (f16 only)"; timeout -k 10 900 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "tmem and not bf16 and not flux" > gpurun_out/r2b_tmem_f16.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2b_tmem_f16.log
echo "== other gpu tests (no tmem)"; timeout -k 10 1500 python -m pytest tests -q -m gpu -k "not tmem and not span and not flux_shape and not sd35 and not per_linear" --deselect tests/test_gpu_linear.py > gpurun_out/r2b_rest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2b_rest.log
