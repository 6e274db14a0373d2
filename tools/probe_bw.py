#!/usr/bin/env python
"""Control experiments for the HBM roofline: what do plain torch kernels reach for copy / write-only / 1B->2B widening
on buffers of the dequant benchmark's size?  (Denominator sanity check for roofline.frac.)"""
import torch

dev = "cuda:0"


def t(fn, n=20, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for elems in (9437184, 66060288, 66060288 * 8):
    srcs = [torch.empty(elems, dtype=torch.float16, device=dev).normal_() for _ in range(3)]
    dsts = [torch.empty(elems, dtype=torch.float16, device=dev) for _ in range(3)]
    u8 = [torch.randint(0, 255, (elems,), dtype=torch.uint8, device=dev) for _ in range(3)]
    st = {"i": 0}

    def nx():
        st["i"] = (st["i"] + 1) % 3
        return st["i"]
    ms = t(lambda: dsts[nx()].copy_(srcs[st["i"]]))
    print(f"elems={elems:10d} copy f16->f16      {ms*1e3:8.1f} us  {elems*4/ms/1e6:8.1f} GB/s")
    ms = t(lambda: dsts[nx()].zero_())
    print(f"elems={elems:10d} memset (write)     {ms*1e3:8.1f} us  {elems*2/ms/1e6:8.1f} GB/s")
    ms = t(lambda: dsts[nx()].copy_(u8[st["i"]]))
    print(f"elems={elems:10d} u8->f16 (1B r,2B w){ms*1e3:8.1f} us  {elems*3/ms/1e6:8.1f} GB/s")
    ms = t(lambda: srcs[nx()].sum())
    print(f"elems={elems:10d} read-only sum      {ms*1e3:8.1f} us  {elems*2/ms/1e6:8.1f} GB/s")
    del srcs, dsts, u8
