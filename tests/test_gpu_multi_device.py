"""Two GPUs in ONE process (ComfyUI moves weights between devices freely, nodes.py:84,117): every kernel must launch on the
device that owns its tensors, and per-device kernel attributes (dynamic shared memory opt-in) must be set on each of them."""
import pytest
import torch
import gguf

import oracle
from util import Q, rel_fro

pytestmark = pytest.mark.gpu


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_gpus(), reason="needs two visible GPUs")
@pytest.mark.parametrize("qt", [Q.Q4_K, Q.Q8_0], ids=lambda q: q.name)
def test_same_results_on_second_device(pkg, qt):
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    N, K = 520, 1024
    raw = torch.from_numpy(oracle.random_blocks(int(qt), N * K // bs, seed=5, scale=0.02).reshape(N, K // bs * ts))
    outs = {}
    for dev in ("cuda:1", "cuda:0", "cuda:1"):          # second device first: its attributes must not depend on device 0
        w = pkg.ops.GGMLTensor(raw.to(dev), tensor_type=qt, tensor_shape=torch.Size((N, K)))
        g = torch.Generator(device="cpu").manual_seed(1)
        x_big = torch.randn(600, K, generator=g).to(dev).to(torch.bfloat16)
        x_small = x_big[:3].contiguous()
        res = [
            pkg.dequant.dequantize_tensor(w, torch.float32),           # > 48 KB of dynamic shared memory
            pkg.dequant.dequantize_tensor(w, torch.float16),
            pkg.ops.linear_packed(x_big, w, None, None, pkg.lib.ALGO_FUSED_MMA),
            pkg.ops.linear_packed(x_big, w, None, None, pkg.lib.ALGO_DEQUANT_MMA),
            pkg.ops.linear_packed(x_big[:64].contiguous(), w, None, None, pkg.lib.ALGO_FUSED_MMA),   # split-K
            pkg.ops.linear_packed(x_small, w, None),                   # GEMV
        ]
        assert all(r.device == torch.device(dev) for r in res)
        outs.setdefault(dev, [r.float().cpu() for r in res])
    for a, b in zip(outs["cuda:0"], outs["cuda:1"]):
        assert torch.equal(a, b)
    W = outs["cuda:0"][0]
    ref = torch.nn.functional.linear(x_big.float().cpu(), W.to(torch.bfloat16).float())
    assert rel_fro(outs["cuda:1"][2].numpy(), ref.numpy()) <= 2e-3
