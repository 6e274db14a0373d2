#!/usr/bin/env python
"""Host-side (Python + launch) cost of one quantised Linear call at a small shape, where the GPU work is far shorter than
the call: us per call for the layer, for ops.linear_packed, for the bare ctypes call, and for torch's F.linear on a dense
weight; then a cProfile of the layer call."""
import cProfile
import os
import pstats
import sys
import time

import torch
import gguf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import oracle  # noqa: E402

ops, lib, dq = ge._sub("ops"), ge._sub("_lib"), ge._sub("dequant")
L = lib.lib()
dev = torch.device("cuda:0")
Q = gguf.GGMLQuantizationType


def wall(fn, n=3000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return dt / n * 1e6


for (M, N, K) in ((16, 512, 512), (512, 512, 512)):
    raw = torch.from_numpy(oracle.random_blocks(int(Q.Q4_K), N * K // 256, seed=0, scale=0.02)).reshape(N, K // 256 * 144).to(dev)
    w = ops.GGMLTensor(raw, tensor_type=Q.Q4_K, tensor_shape=torch.Size((N, K)))
    lin = ops.GGMLOps.Linear(K, N)
    lin.load_state_dict({"weight": w})
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    dense = dq.dequantize_tensor(w, torch.bfloat16).as_subclass(torch.Tensor)
    ws = torch.empty(max(1, L.ggufb200_linear_workspace(int(Q.Q4_K), M, N, K, 1, 0)), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def raw_call():
        L.ggufb200_linear(int(Q.Q4_K), raw.data_ptr(), N, K, x.data_ptr(), M, K, 1, 0, None, 0, y.data_ptr(), N, ws.data_ptr(), ws.numel(), 0, st)

    print(f"M={M} N={N} K={K}: layer {wall(lambda: lin(x)):.1f} us | linear_packed {wall(lambda: ops.linear_packed(x, w, None)):.1f} us | "
          f"bare ctypes call {wall(raw_call):.1f} us | F.linear(dense) {wall(lambda: torch.nn.functional.linear(x, dense)):.1f} us", flush=True)

pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    lin(x)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
