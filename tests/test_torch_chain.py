"""CPU test (build container only): oracle/torch_chain.py reproduces the reference's torch path bit-for-bit and with
the same number of materialising ATen kernels, so that timing it on the GPU box is a fair stand-in for the reference."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import gguf
from torch.utils._python_dispatch import TorchDispatchMode

import oracle
from oracle import torch_chain

REF = "/root/reference/dequant.py"
Q = gguf.GGMLQuantizationType
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")


class Count(TorchDispatchMode):
    VIEWS = ("view", "reshape", "split", "unsqueeze", "expand", "_unsafe_view", "alias", "detach", "select", "slice", "squeeze",
             "permute", "transpose", "t.default", "split_with_sizes", "_reshape_alias", "lift_fresh")

    def __init__(self):
        super().__init__()
        self.n = 0

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not any(v in name for v in self.VIEWS):
            self.n += 1
        return func(*args, **(kwargs or {}))


def _ref():
    spec = importlib.util.spec_from_file_location("ref_dequant", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("qt", [Q.Q8_0, Q.Q4_0, Q.Q4_K, Q.Q5_K, Q.Q6_K, Q.BF16], ids=lambda q: q.name)
@pytest.mark.parametrize("md", [None, torch.float32, torch.bfloat16])
def test_chain_equals_reference(qt, md):
    ref = _ref()
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    raw = torch.from_numpy(oracle.random_blocks(int(qt), 96, seed=1)).reshape(-1)
    shape = (96 * bs,)
    with Count() as c_ref:
        a = ref.dequantize(raw, qt, shape, dtype=md)
    with Count() as c_mine:
        b = torch_chain.dequantize(raw, qt, shape, dtype=md)
    assert a.dtype == b.dtype and torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    assert c_ref.n == c_mine.n, (c_ref.n, c_mine.n)
