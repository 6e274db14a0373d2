"""GPU test: a shallow Flux-shape DiT built through the drop-in GGMLOps gives the same output as the same network run
through the reference's torch chain (oracle/torch_chain.py, pinned to the reference) on identical packed weights."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("numerics", ["exact", "fast"])
def test_flux_shape_block_matches_reference_chain(pkg, numerics):
    import flux_harness as fh
    cls = pkg.ops.GGMLOps.Linear
    before = cls.linear_numerics
    cls.linear_numerics = numerics
    try:
        dev = torch.device("cuda:0")
        with torch.no_grad():
            ours = fh.FluxShapeDiT(pkg.ops.GGMLOps, depth=1, depth_single=1)
            ref = fh.FluxShapeDiT(fh.RefChainOps, depth=1, depth_single=1)
            sd = fh.build_state_dict(ours, pkg.ops.GGMLTensor, dev)
            fh.load_shared(ours, sd).to(dev)
            fh.load_shared(ref, sd).to(dev)
            inp = fh.make_inputs(dev, torch.bfloat16, img_tokens=1024, txt_tokens=256)
            a, b = ours(**inp), ref(**inp)
        assert type(a) is torch.Tensor and a.shape == b.shape == (1, 1024, 64)
        assert torch.isfinite(a).all()
        rel = ((a.float() - b.float()).norm() / b.float().norm()).item()
        assert rel <= 1e-2, rel      # whole-network drift through two blocks of bf16 ops; per-Linear parity is 1e-3 (test_gpu_gemm)
    finally:
        cls.linear_numerics = before


@pytest.mark.parametrize("numerics", ["exact", "fast"])
def test_per_linear_parity_at_flux_scale(pkg, numerics):
    """BASELINE config 3 scale (hidden 3072, 4096 + 512 tokens, Q4_K block Linears, bf16), two double + two single blocks:
    EVERY quantised Linear, on the route AUTO picks for it, against the reference arithmetic on the same input (bit-exact K1
    weight, fp32 accumulate).  exact: <= 1e-3 (north_star).  fast: <= 8e-3 for bf16 activations."""
    import flux_harness as fh
    cls = pkg.ops.GGMLOps.Linear
    before = cls.linear_numerics
    cls.linear_numerics = numerics
    try:
        dev = torch.device("cuda:0")
        with torch.no_grad():
            model = fh.FluxShapeDiT(pkg.ops.GGMLOps, depth=2, depth_single=2)
            fh.load_shared(model, fh.build_state_dict(model, pkg.ops.GGMLTensor, dev)).to(dev)
            inp = fh.make_inputs(dev, torch.bfloat16)
            with fh.LinearParity(model, pkg.dequant) as lp:
                model(**inp)
        assert lp.records and len(lp.records) >= 2 * 10 + 2 * 3
        budget = 1e-3 if numerics == "exact" else 8e-3
        worst = max(lp.records, key=lambda r: r[-1])
        assert worst[-1] <= budget, worst
        assert {r[2] for r in lp.records} >= {1, 512, 4096, 4608}      # GEMV, short and long activations all covered
    finally:
        cls.linear_numerics = before
