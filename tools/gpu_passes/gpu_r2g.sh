#!/bin/bash
# round-2 GPU pass G: full suite with the new AUTO routing, bench.py, Flux-shape step, per-Linear parity, ncu text exports
set +e
mkdir -p gpurun_out
for f in test_gpu_dequant test_gpu_gemm test_gpu_linear test_gpu_flux; do
  timeout -k 10 1500 python -m pytest tests/$f.py -q -m gpu > gpurun_out/r2g_$f.log 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/r2g_$f.log | head -3
  grep -E "^FAILED" gpurun_out/r2g_$f.log | head -12
done
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench.py"; timeout -k 10 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2g_bench.json") if l.startswith("{")][-1])
    print("value",d["value"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"] if d.get("e2e") else None)
    print("per_qtype",{k:round(v["frac"],3) for k,v in d["roofline"]["per_qtype"].items()})
    f=d["flux_step"]; print("flux",{k:f.get(k) for k in ("ms_per_step","reference_chain_ms_per_step","linear_ms","other_ms","output_rel_err_vs_reference_chain","numerics","error")})
    print("cpu",d["cpu_baseline"]["value"],d["cpu_baseline"].get("numpy_gguf_py",{}).get("value"))
except Exception as e: print("bench parse failed",e)
PY
echo "== flux fast numerics"; timeout -k 10 600 python tools/bench_flux.py --numerics fast --ref-steps 0 > gpurun_out/r2g_flux_fast.json 2>gpurun_out/r2g_flux_fast.err; cut -c1-600 gpurun_out/r2g_flux_fast.json
echo "== per-Linear parity flux"; timeout -k 10 900 python tools/bench_models.py parity flux > gpurun_out/r2g_parity_flux.json 2>gpurun_out/r2g_parity.err; cut -c1-900 gpurun_out/r2g_parity_flux.json; tail -3 gpurun_out/r2g_parity.err
echo "== ncu gemm4 tile384 v3"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm4_kernel -s 3 -c 1 -f -o /tmp/g4a python tools/bench_linear.py --M 4608 --shapes 2 --routes tmem384 > gpurun_out/r2g_ncu1.log 2>&1; tail -1 gpurun_out/r2g_ncu1.log
ncu -i /tmp/g4a.ncu-rep --page raw --csv > gpurun_out/r02_gemm4_v3_tile384_m4608_raw.csv 2>/dev/null
ncu -i /tmp/g4a.ncu-rep --page details > gpurun_out/r02_gemm4_v3_tile384_m4608_details.txt 2>/dev/null
echo "== ncu gemm3 dense"
timeout -k 10 600 ncu --set full --clock-control none -k regex:gemm3_kernel -s 3 -c 1 -f -o /tmp/g3 python tools/bench_linear.py --M 4608 --shapes 2 --routes ours_dense > gpurun_out/r2g_ncu2.log 2>&1; tail -1 gpurun_out/r2g_ncu2.log
ncu -i /tmp/g3.ncu-rep --page raw --csv > gpurun_out/r02_gemm3_v2_dense_m4608_raw.csv 2>/dev/null
ncu -i /tmp/g3.ncu-rep --page details > gpurun_out/r02_gemm3_v2_dense_m4608_details.txt 2>/dev/null
du -sh gpurun_out
