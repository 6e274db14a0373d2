#!/usr/bin/env python
"""Kernel-level breakdown of one Flux-shape step of this repo's GGMLOps (torch.profiler, CUDA time by kernel name)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
import flux_harness as fh  # noqa: E402

ops_mod = ge._sub("ops")
dev = torch.device("cuda:0")
with torch.no_grad():
    model = fh.FluxShapeDiT(ops_mod.GGMLOps)
    fh.load_shared(model, fh.build_state_dict(model, ops_mod.GGMLTensor, dev)).to(dev)
    inp = fh.make_inputs(dev, torch.bfloat16)
    for _ in range(2):
        model(**inp)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(**inp)
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows if not r[0].startswith("aten::") and not r[0].startswith("cuda"))
print("kernels by GPU time (us), one step")
for k, t, c in rows[:40]:
    print(f"{t:10.0f} us  x{c:4d}  {k[:150]}")
