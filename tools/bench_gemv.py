#!/usr/bin/env python
"""GPU-side time of the small-M fused Linear (K3): raw C-ABI launches (no Python wrapper work in the loop), weights rotated
through > L2 worth of buffers, both GEMV kernels.  Prints GB/s of packed weight read vs the HBM peak."""
import json
import os
import sys

import torch
import gguf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import oracle  # noqa: E402

ops, lib = ge._sub("ops"), ge._sub("_lib")
L = lib.lib()
dev = torch.device("cuda:0")
peak = 6574.5
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
st = torch.cuda.current_stream().cuda_stream
for qname in (sys.argv[1:] or ["Q4_K", "Q8_0"]):
    qt = gguf.GGMLQuantizationType[qname]
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    for (N, K) in ((18432, 3072), (9216, 3072)):
        copies = 6
        ws = []
        for c in range(copies):
            raw = torch.from_numpy(oracle.random_blocks(int(qt), 1 << 14, seed=c, scale=0.02))
            reps = (N * K // bs + (1 << 14) - 1) // (1 << 14)
            ws.append(raw.repeat(reps, 1)[: N * K // bs].reshape(N, K // bs * ts).contiguous().to(dev))
        for M in (1, 4, 8):
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for variant in (1, 0):
                L.ggufb200_set_tuning(5, variant)

                def launch(i):
                    rc = L.ggufb200_linear(int(qt), ws[i % copies].data_ptr(), N, K, x.data_ptr(), M, K, 1, 0, None, 0, y.data_ptr(), N, None, 0, 1, st)
                    assert rc == 0, rc
                for i in range(10):
                    launch(i)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 120
                a.record()
                for i in range(iters):
                    launch(i)
                b.record()
                torch.cuda.synchronize()
                us = a.elapsed_time(b) / iters * 1e3
                gbs = N * K // bs * ts / us / 1e3
                print(f"{qname} N={N} K={K} M={M} {'mma' if variant else 'fma'}: {us:7.1f} us  {gbs:7.1f} GB/s packed read  ({gbs / peak:.3f} of HBM peak)", flush=True)
L.ggufb200_set_tuning(5, 1)
