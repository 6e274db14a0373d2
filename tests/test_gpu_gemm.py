"""GPU parity tests of the tcgen05 GEMM kernel (csrc/gemm.cu): dense TMA-fed mode and fused-dequant mode.

Reference for both: fp32-accumulated x @ W^T (+bias) rounded to the activation dtype, where W is the bit-exact
dequantised weight (validated separately against the reference).  Tolerance 1e-3 relative (Frobenius)."""
import numpy as np
import pytest
import torch
import gguf

import oracle
from util import Q, rel_fro

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3


@pytest.fixture(params=[(2, 1), (1, 1), (1, 0), (0, 0)], ids=["persistent+pair-staged", "pair-staged", "pair", "single"], autouse=True)
def gemm_variant(request, pkg):
    """Every test runs on the CTA-pair kernel (cta_group::2, default) with and without TMA-staged packed tiles in the
    fused producer, and on the single-CTA kernel."""
    variant, staged = request.param
    pkg.lib.lib().ggufb200_set_tuning(2, variant)
    pkg.lib.lib().ggufb200_set_tuning(4, staged)
    yield request.param
    pkg.lib.lib().ggufb200_set_tuning(2, 2)
    pkg.lib.lib().ggufb200_set_tuning(4, 1)


def _ref(x, W, bias):
    y = x.float() @ W.float().t()
    if bias is not None:
        y = y + bias.to(x.dtype).float()
    return y.to(x.dtype)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 264, 512), (1000, 128, 3072), (24, 512, 256), (513, 1032, 1024)])
def test_dense_gemm_matches_torch(pkg, dt, M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn(M, K, device=DEV, dtype=dt, generator=g)
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(dt)
    b = torch.randn(N, device=DEV, generator=g) * 0.1
    y = pkg.ops.linear_dense(x, W, b)
    ref = _ref(x, W, b)
    assert rel_fro(y.float().cpu().numpy(), ref.float().cpu().numpy()) <= TOL
    y2 = pkg.ops.linear_dense(x, W, None)
    assert rel_fro(y2.float().cpu().numpy(), _ref(x, W, None).float().cpu().numpy()) <= TOL


@pytest.mark.parametrize("qt", [Q.Q4_0, Q.Q4_1, Q.Q5_0, Q.Q5_1, Q.Q8_0, Q.Q2_K, Q.Q3_K, Q.Q4_K, Q.Q5_K, Q.Q6_K, Q.IQ4_NL, Q.IQ4_XS],
                         ids=lambda q: q.name)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_fused_gemm_all_types(pkg, qt, dt):
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    M, N, K = 300, 264, 1024
    raw = oracle.random_blocks(int(qt), N * K // bs, seed=int(qt), scale=0.02).reshape(N, K // bs * ts)
    w = pkg.ops.GGMLTensor(torch.from_numpy(raw).to(DEV), tensor_type=qt, tensor_shape=torch.Size((N, K)))
    x = torch.randn(M, K, device=DEV, dtype=dt)
    b = torch.randn(N, device=DEV) * 0.1
    y = pkg.ops.linear_packed(x, w, b, None, pkg.lib.ALGO_FUSED_MMA)
    W = pkg.dequant.dequantize_tensor(w, dt)
    assert rel_fro(y.float().cpu().numpy(), _ref(x, W, b).float().cpu().numpy()) <= TOL


@pytest.mark.parametrize("qt", [Q.Q4_K, Q.Q8_0, Q.Q6_K, Q.Q5_0], ids=lambda q: q.name)
@pytest.mark.parametrize("M,N,K", [(64, 512, 4096), (300, 264, 2048), (513, 520, 1280), (1000, 256, 5120), (24, 1032, 768)])
def test_fused_gemm_split_k(pkg, gemm_variant, qt, M, N, K):
    """Short activations: the fused kernel cuts K into ranges of whole 256-wide spans, one SM pair per (tile, range), and
    keeps fp32 partial tiles in the workspace and sums them in a fixed order (bit-reproducible).  Checked against the reference arithmetic and against the unsplit kernel
    (same tiles, no workspace): the two may differ only by fp32 summation order."""
    L = pkg.lib.lib()
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    raw = oracle.random_blocks(int(qt), N * K // bs, seed=int(qt) + K, scale=0.02).reshape(N, K // bs * ts)
    w = pkg.ops.GGMLTensor(torch.from_numpy(raw).to(DEV), tensor_type=qt, tensor_shape=torch.Size((N, K)))
    W = pkg.dequant.dequantize_tensor(w, torch.bfloat16)
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, device=DEV) * 0.1
    variant, _staged = gemm_variant
    L.ggufb200_set_tuning(6, 1)
    need = L.ggufb200_linear_workspace(int(qt), M, N, K, 1, pkg.lib.ALGO_FUSED_MMA)
    if variant == 0:
        assert need == 0                    # the single-CTA kernel has no split-K
    else:
        assert need % (M * N * 4) == 0 and need >= 2 * M * N * 4, "these shapes leave SM pairs idle without split-K"
    y = pkg.ops.linear_packed(x, w, b, None, pkg.lib.ALGO_FUSED_MMA)
    assert torch.equal(y, pkg.ops.linear_packed(x, w, b, None, pkg.lib.ALGO_FUSED_MMA)), "split-K must be reproducible"
    L.ggufb200_set_tuning(6, 0)
    try:
        assert L.ggufb200_linear_workspace(int(qt), M, N, K, 1, pkg.lib.ALGO_FUSED_MMA) == 0
        y1 = pkg.ops.linear_packed(x, w, b, None, pkg.lib.ALGO_FUSED_MMA)
    finally:
        L.ggufb200_set_tuning(6, 1)
    ref = _ref(x, W, b).float().cpu().numpy()
    assert rel_fro(y.float().cpu().numpy(), ref) <= TOL
    assert rel_fro(y1.float().cpu().numpy(), ref) <= TOL
    assert rel_fro(y.float().cpu().numpy(), y1.float().cpu().numpy()) <= 2e-3   # bf16 output rounding of two fp32 orders
    ya = pkg.ops.linear_packed(x, w, None, None, pkg.lib.ALGO_AUTO)
    assert rel_fro(ya.float().cpu().numpy(), _ref(x, W, None).float().cpu().numpy()) <= TOL


def test_fused_split_k_without_workspace_runs_unsplit(pkg):
    """The C ABI never allocates: FUSED_MMA with a NULL workspace must still be correct (one pair per tile)."""
    qt, M, N, K = Q.Q4_K, 64, 512, 4096
    raw = oracle.random_blocks(int(qt), N * K // 256, seed=9, scale=0.02).reshape(N, K // 256 * 144)
    w = torch.from_numpy(raw).to(DEV)
    x = torch.randn(M, K, device=DEV, dtype=torch.float16)
    y = torch.empty(M, N, device=DEV, dtype=torch.float16)
    L = pkg.lib.lib()
    rc = L.ggufb200_linear(int(qt), w.data_ptr(), N, K, x.data_ptr(), M, K, 0, 0, None, 0, y.data_ptr(), N, None, 0,
                           pkg.lib.ALGO_FUSED_MMA, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    gw = pkg.ops.GGMLTensor(w, tensor_type=qt, tensor_shape=torch.Size((N, K)))
    W = pkg.dequant.dequantize_tensor(gw, torch.float16)
    assert rel_fro(y.float().cpu().numpy(), _ref(x, W, None).float().cpu().numpy()) <= TOL


@pytest.mark.parametrize("route", ["fused", "dequant+mma", "auto"])
def test_sd35_shape_q8_0_unaligned_rows(pkg, route):
    """SD3.5-large hidden size 2432: Q8_0 rows are 2584 bytes (not a multiple of 16), so no tensor map over the packed bytes
    is legal; the fused kernel must take its direct-load producer and K1 its flat byte-stream tiling."""
    M, N, K = 700, 7296, 2432
    raw = oracle.random_blocks(int(Q.Q8_0), N * K // 32, seed=2, scale=0.02).reshape(N, K // 32 * 34)
    w = pkg.ops.GGMLTensor(torch.from_numpy(raw).to(DEV), tensor_type=Q.Q8_0, tensor_shape=torch.Size((N, K)))
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, device=DEV) * 0.1
    algo = {"fused": pkg.lib.ALGO_FUSED_MMA, "dequant+mma": pkg.lib.ALGO_DEQUANT_MMA, "auto": pkg.lib.ALGO_AUTO}[route]
    y = pkg.ops.linear_packed(x, w, b, None, algo)
    W = pkg.dequant.dequantize_tensor(w, torch.bfloat16)
    assert rel_fro(y.float().cpu().numpy(), _ref(x, W, b).float().cpu().numpy()) <= TOL


@pytest.mark.parametrize("M,N,K", [(4608, 3072, 3072), (512, 9216, 3072), (4096, 3072, 12288)])
def test_fused_gemm_flux_shapes_q4k(pkg, M, N, K):
    """Full Flux.1 Linear sizes: the oracle is too slow here, so compare against the dense route on the same weight
    (K1 dequant, bit-exact vs the reference, then a library fp32-accumulate matmul)."""
    qt = Q.Q4_K
    raw = oracle.random_blocks(int(qt), N * K // 256, seed=1, scale=0.02).reshape(N, K // 256 * 144)
    w = pkg.ops.GGMLTensor(torch.from_numpy(raw).to(DEV), tensor_type=qt, tensor_shape=torch.Size((N, K)))
    x = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    y = pkg.ops.linear_packed(x, w, None, None, pkg.lib.ALGO_FUSED_MMA)
    W = pkg.dequant.dequantize_tensor(w, torch.bfloat16)
    ref = torch.nn.functional.linear(x, W)
    assert rel_fro(y.float().cpu().numpy(), ref.float().cpu().numpy()) <= TOL
    y3 = pkg.ops.linear_packed(x, w, None, None, pkg.lib.ALGO_DEQUANT_MMA)
    assert rel_fro(y3.float().cpu().numpy(), ref.float().cpu().numpy()) <= TOL


def test_auto_route_large_m_through_layer(pkg):
    lin = pkg.ops.GGMLOps.Linear(1024, 264)
    raw = oracle.random_blocks(int(Q.Q5_K), 264 * 4, seed=3, scale=0.02).reshape(264, 4 * 176)
    w = pkg.ops.GGMLTensor(torch.from_numpy(raw).to(DEV), tensor_type=Q.Q5_K, tensor_shape=torch.Size((264, 1024)))
    lin.load_state_dict({"weight": w})
    x = torch.randn(2, 200, 1024, device=DEV, dtype=torch.bfloat16)
    y = lin(x)
    W = pkg.dequant.dequantize_tensor(w, torch.bfloat16)
    assert tuple(y.shape) == (2, 200, 264)
    assert rel_fro(y.float().cpu().numpy(), _ref(x.reshape(-1, 1024), W, None).float().cpu().numpy()) <= TOL
