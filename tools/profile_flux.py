#!/usr/bin/env python
"""Kernel-time breakdown of one Flux-shape step (torch profiler, CUDA activities) for either arm."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as ge  # noqa: E402
import flux_harness as fh  # noqa: E402

arm = sys.argv[1] if len(sys.argv) > 1 else "ours"
ops_mod = ge._sub("ops")
dev = torch.device("cuda:0")
with torch.no_grad():
    model = fh.FluxShapeDiT(ops_mod.GGMLOps if arm == "ours" else fh.RefChainOps)
    sd = fh.build_state_dict(model, ops_mod.GGMLTensor, dev)
    fh.load_shared(model, sd).to(dev)
    inp = fh.make_inputs(dev, torch.bfloat16)
    for _ in range(2):
        model(**inp)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(**inp)
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"arm={arm} total device time {tot:.1f} ms")
for k, ms, n in rows[:28]:
    print(f"{ms:9.2f} ms  {100 * ms / tot:5.1f}%  x{n:<5d} {k[:110]}")
