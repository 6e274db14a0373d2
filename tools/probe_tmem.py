#!/usr/bin/env python
"""One gemm4 case per process (a device-side fault poisons the CUDA context): python tools/probe_tmem.py <act> <flags-hex> [M N K qtype]."""
import os
import sys

import torch
import gguf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import oracle  # noqa: E402

ops, dq, lib = ge._sub("ops"), ge._sub("dequant"), ge._sub("_lib")
act = torch.float16 if sys.argv[1] == "f16" else torch.bfloat16
flags = int(sys.argv[2], 16)
M, N, K = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (300, 264, 1024)
qt = gguf.GGMLQuantizationType[sys.argv[6]] if len(sys.argv) > 6 else gguf.GGMLQuantizationType.Q4_K
bs, ts = gguf.GGML_QUANT_SIZES[qt]
dev = torch.device("cuda:0")
raw = oracle.random_blocks(int(qt), N * K // bs, seed=3, scale=0.02).reshape(N, K // bs * ts)
w = ops.GGMLTensor(torch.from_numpy(raw).to(dev), tensor_type=qt, tensor_shape=torch.Size((N, K)))
x = torch.randn(M, K, device=dev, dtype=act)
W = dq.dequantize_tensor(w, act)
ref = (x.float() @ W.float().t())
torch.cuda.synchronize()
y = ops.linear_packed(x, w, None, None, lib.ALGO_FUSED_TMEM | flags)
torch.cuda.synchronize()
err = float(((y.float() - ref).norm() / ref.norm()).item())
print(f"probe act={sys.argv[1]} flags={flags:#x} M={M} N={N} K={K} {qt.name}: relerr={err:.3e} finite={bool(torch.isfinite(y).all())}", flush=True)
