#!/usr/bin/env python
"""Dense tcgen05 GEMM: time vs K at fixed M, N to split fixed per-tile overhead from per-k-block cost; ours vs cuBLAS."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ops, lib = ge._sub("ops"), ge._sub("_lib")
dev = torch.device("cuda:0")


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for variant in (2, 1):
    lib.lib().ggufb200_set_tuning(2, variant)
    for (M, N) in ((4608, 12288), (4736, 9472)):     # 2nd: 18.5x37 = exactly full waves of 512x256 pair tiles? (tiles=10*37=370=5*74)
        res = []
        for K in (512, 1024, 2048, 3072, 6144, 12288):
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(2)]
            i = [0]

            def ours():
                i[0] ^= 1
                return ops.linear_dense(x, ws[i[0]])

            def cub():
                i[0] ^= 1
                return torch.nn.functional.linear(x, ws[i[0]])
            t1, t2 = timeit(ours), timeit(cub)
            fl = 2.0 * M * N * K
            res.append((K, t1, t2))
            print(f"variant={variant} M={M} N={N} K={K:6d} ours {t1*1e3:8.1f} us {fl/t1/1e9:7.1f} TF | cublas {t2*1e3:8.1f} us {fl/t2/1e9:7.1f} TF", flush=True)
        ks = np.array([r[0] for r in res], float)
        for name, col in (("ours", 1), ("cublas", 2)):
            ts = np.array([r[col] for r in res]) * 1e3
            b, a = np.polyfit(ks, ts, 1)
            print(f"   fit {name}: t = {a:.1f} us + {b*64:.3f} us per k-block(64)  -> asymptotic {2.0*M*N*64/(b*64)/1e6:.0f} TF")
