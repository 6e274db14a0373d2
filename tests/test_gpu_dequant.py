"""GPU parity tests of the standalone dequant kernel (csrc/dequant.cu) through the C ABI.

Bar: BIT-EXACT against (a) the golden vectors produced by the unmodified reference and (b) the CPU oracle,
for every supported type and every (math dtype, out dtype) pair -- not a tolerance."""
import os

import numpy as np
import pytest
import torch
import gguf

import oracle
from util import ALL_QTYPES, COMBOS, Q, TORCH_DT, canon_nan, torch_bits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gt(pkg, raw, qt, shape):
    return pkg.ops.GGMLTensor(torch.from_numpy(np.ascontiguousarray(raw)).to(DEV), tensor_type=qt, tensor_shape=torch.Size(shape))


@pytest.mark.parametrize("qt", ALL_QTYPES, ids=lambda q: q.name)
def test_dequant_matches_reference_golden_bit_exact(pkg, qt, golden_dir):
    g = np.load(os.path.join(golden_dir, f"dequant_{qt.name}.npz"))
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = int(g["n_blocks"])
    packed = torch.from_numpy(g["packed"]).to(DEV)
    for math, od in COMBOS:
        md = None if math == 0 else TORCH_DT[math]
        out = pkg.dequant.dequantize(packed, qt, (n_blocks * bs,), dtype=md, out_dtype=TORCH_DT[od])
        assert out.dtype == TORCH_DT[od]
        got = canon_nan(torch_bits(out), od)
        want = canon_nan(g[f"out_m{math}_o{od}"], od)
        assert np.array_equal(got, want), f"{qt.name} math={math} out={od}: {np.count_nonzero(got != want)} mismatches"


@pytest.mark.parametrize("qt", ALL_QTYPES, ids=lambda q: q.name)
@pytest.mark.parametrize("math,od", [(0, 0), (0, 1), (2, 2), (1, 1), (1, 0), (2, 1)])
def test_dequant_matches_oracle_ragged_sizes(pkg, qt, math, od):
    """Block counts that are not tile multiples (tile = 2048 elements), one-block and prime-sized inputs."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    sizes = (1, 7, 8, 9, 1031) if bs == 256 else ((1, 63, 64, 65, 4099) if bs == 32 else (1, 7, 2049, 40001))
    for n_blocks in sizes:
        raw = oracle.random_blocks(int(qt), n_blocks, seed=n_blocks)
        md = None if math == 0 else TORCH_DT[math]
        out = pkg.dequant.dequantize(torch.from_numpy(raw).to(DEV), qt, (n_blocks * bs,), dtype=md, out_dtype=TORCH_DT[od])
        want = oracle.dequant(raw, int(qt), od, math)
        want = want.view(np.uint32) if od == 2 else want
        assert np.array_equal(canon_nan(torch_bits(out), od), canon_nan(want, od)), (qt.name, n_blocks)


@pytest.mark.parametrize("qt", [q for q in ALL_QTYPES if q != Q.BF16], ids=lambda q: q.name)
def test_integer_unpack_bit_exact(pkg, qt):
    raw = oracle.random_blocks(int(qt), 300, seed=3)
    raw[0, :] = 0
    raw[1, :] = 0xFF
    q, sc, mn = pkg.dequant.unpack_int(torch.from_numpy(raw).to(DEV), qt)
    oq, osc, omn = oracle.unpack_int(raw, int(qt))
    assert np.array_equal(q.cpu().numpy(), oq)
    assert np.array_equal(sc.cpu().numpy(), osc)
    assert np.array_equal(mn.cpu().numpy(), omn)


@pytest.mark.parametrize("qt", [Q.Q4_0, Q.Q4_K, Q.Q6_K, Q.Q8_0, Q.BF16], ids=lambda q: q.name)
def test_unaligned_source_pointer(pkg, qt):
    """A packed payload that does not start on a 16-byte boundary (a sliced view) takes the non-TMA staging path."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = 531 if bs > 1 else 5001
    raw = oracle.random_blocks(int(qt), n_blocks, seed=9).reshape(-1)
    buf = torch.zeros(raw.size + 64, dtype=torch.uint8, device=DEV)
    for shift in (2, 6):
        buf[shift:shift + raw.size] = torch.from_numpy(raw).to(DEV)
        view = buf[shift:shift + raw.size]
        assert view.data_ptr() % 16 != 0
        out = pkg.dequant.dequantize(view, qt, (n_blocks * bs,), dtype=None, out_dtype=torch.float16)
        want = oracle.dequant(raw, int(qt), 0, 0)
        assert np.array_equal(torch_bits(out), want)


def test_empty_input(pkg):
    out = pkg.dequant.dequantize(torch.zeros(0, dtype=torch.uint8, device=DEV), Q.Q4_K, (0,), out_dtype=torch.float16)
    assert out.numel() == 0


def test_dequantize_tensor_semantics(pkg):
    """dtype=None keeps the math dtype (fp16; fp32 for BF16), 'target' runs the math in the requested dtype, CPU tensors
    are dequantised on the GPU and returned on the CPU (no CPU compute path exists)."""
    raw = oracle.random_blocks(int(Q.Q4_K), 24, seed=1).reshape(3, 8 * 144)
    t = _gt(pkg, raw, Q.Q4_K, (3, 2048))
    a = pkg.dequant.dequantize_tensor(t)
    assert a.dtype == torch.float16 and tuple(a.shape) == (3, 2048)
    b = pkg.dequant.dequantize_tensor(t, torch.bfloat16, "target")
    assert np.array_equal(torch_bits(b), oracle.dequant(raw, int(Q.Q4_K), 1, 1))
    c = pkg.dequant.dequantize_tensor(t, torch.float32, torch.float32)
    assert np.array_equal(torch_bits(c), oracle.dequant(raw, int(Q.Q4_K), 2, 2).view(np.uint32))
    cpu_t = pkg.ops.GGMLTensor(torch.from_numpy(raw), tensor_type=Q.Q4_K, tensor_shape=torch.Size((3, 2048)))
    d = pkg.dequant.dequantize_tensor(cpu_t, torch.float16)
    assert d.device.type == "cpu" and np.array_equal(torch_bits(d), oracle.dequant(raw, int(Q.Q4_K), 0, 0))
    rawb = oracle.random_blocks(30, 64, seed=2)
    e = pkg.dequant.dequantize_tensor(_gt(pkg, rawb.reshape(-1), Q.BF16, (64,)))
    assert e.dtype == torch.float32                         # dequant.py:61-62


@pytest.mark.parametrize("qt,shape", [(Q.Q4_K, (21504, 3072)), (Q.Q8_0, (3072, 12288)), (Q.Q4_0, (9216, 3072)), (Q.Q6_K, (3072, 3072))],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_full_size_flux_shapes_bit_exact(pkg, qt, shape):
    """BASELINE config-2 sizes.  The C oracle is fast enough to check every element; plus idempotence of repeat launches."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    N, K = shape
    raw = oracle.random_blocks(int(qt), N * K // bs, seed=123)
    t = _gt(pkg, raw.reshape(N, K // bs * ts), qt, shape)
    out1 = pkg.dequant.dequantize_tensor(t, torch.float16)
    out2 = pkg.dequant.dequantize_tensor(t, torch.float16)
    assert torch.equal(out1, out2)
    want = oracle.dequant(raw, int(qt), 0, 0)
    assert np.array_equal(torch_bits(out1), want)


@pytest.mark.parametrize("qt", [Q.Q4_K, Q.Q5_K, Q.Q8_0, Q.Q6_K, Q.Q4_0, Q.BF16], ids=lambda q: q.name)
def test_row_gather_matches_full_dequant(pkg, qt):
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    N, K = 300, 4096 if bs > 1 else 520
    raw = oracle.random_blocks(int(qt), N * K // bs, seed=4).reshape(N, K // bs * ts)
    t = _gt(pkg, raw, qt, (N, K))
    idx = torch.tensor([[0, 299, 17], [17, 5, 128]], device=DEV)
    got = pkg.dequant.dequantize_rows(t, idx, torch.bfloat16)
    full = pkg.dequant.dequantize_tensor(t, torch.bfloat16)
    assert tuple(got.shape) == (2, 3, K)
    assert torch.equal(got, full[idx])


@pytest.mark.parametrize("qt", [Q.Q4_0, Q.Q8_0, Q.Q4_K, Q.Q6_K, Q.Q3_K, Q.IQ4_XS], ids=lambda q: q.name)
def test_src_stable_flag_changes_nothing_but_the_launch(pkg, qt):
    """GGUFB200_DEQUANT_SRC_STABLE (include/ggufb200.h) only moves the griddepcontrol.wait: same bits with and without it,
    also for short last tiles and for a 16-byte-misaligned source (the flag is ignored there)."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    for n_blocks in ((5, 16 * 7 + 3) if bs == 256 else (130, 128 * 9 + 77)):
        raw = oracle.random_blocks(int(qt), n_blocks, seed=11 + n_blocks)
        want = oracle.dequant(raw, int(qt), oracle.DT_F16, oracle.DT_F16)
        dev_raw = torch.from_numpy(raw).to(DEV)
        for stable in (False, True):
            out = pkg.dequant.dequantize(dev_raw, qt, (n_blocks * bs,), src_stable=stable)
            assert np.array_equal(canon_nan(torch_bits(out), 0), canon_nan(want, 0)), (qt.name, n_blocks, stable)
        shifted = torch.empty(raw.size + 2, dtype=torch.uint8, device=DEV)[2:]
        shifted.copy_(dev_raw.reshape(-1))
        out = pkg.dequant.dequantize(shifted, qt, (n_blocks * bs,), src_stable=True)
        assert np.array_equal(canon_nan(torch_bits(out), 0), canon_nan(want, 0)), (qt.name, n_blocks, "misaligned")


def test_back_to_back_launches_into_one_buffer_keep_stream_order(pkg):
    """Programmatic dependent launch lets a dequant kernel start (fetch + unpack) under its predecessor's tail; its STORES must
    still wait.  A CUDA graph of 24 raw launches alternates two different packed tensors into ONE output buffer, with a
    device-side copy of the buffer after every launch: every copy must hold exactly the tensor launched just before it."""
    L, lib = pkg.lib.lib(), pkg.lib
    qt = Q.Q4_K
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = 16 * 600 + 5                   # 601 tiles: several waves of CTAs, short last tile
    raws = [oracle.random_blocks(int(qt), n_blocks, seed=s) for s in (1, 2)]
    wants = [oracle.dequant(r, int(qt), oracle.DT_F16, oracle.DT_F16) for r in raws]
    devs = [torch.from_numpy(r).to(DEV) for r in raws]
    out = torch.zeros(n_blocks * bs, dtype=torch.float16, device=DEV)
    snaps = [torch.empty_like(out) for _ in range(24)]
    side = torch.cuda.Stream()
    torch.cuda.synchronize()

    def enqueue(st):
        for i in range(24):
            rc = L.ggufb200_dequant(int(qt), devs[i & 1].data_ptr(), n_blocks, out.data_ptr(), 0, lib.DEQUANT_SRC_STABLE, st.cuda_stream)
            assert rc == 0
            with torch.cuda.stream(st):
                snaps[i].copy_(out)
    with torch.cuda.stream(side):
        enqueue(side)           # eager once (also warms the kernels before capture)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        enqueue(side)
    for s in snaps:
        s.zero_()
    g.replay()
    torch.cuda.synchronize()
    for i, s in enumerate(snaps):
        assert np.array_equal(torch_bits(s), wants[i & 1]), f"launch {i}: the buffer does not hold the tensor launched last"


def test_consecutive_dequants_without_a_kernel_between_them(pkg):
    """Same as above with NOTHING between the launches (kernel -> kernel programmatic edges only): the last writer wins."""
    L, lib = pkg.lib.lib(), pkg.lib
    qt = Q.Q8_0
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = 128 * 700 + 9
    raws = [oracle.random_blocks(int(qt), n_blocks, seed=s) for s in (5, 6, 7)]
    devs = [torch.from_numpy(r).to(DEV) for r in raws]
    out = torch.zeros(n_blocks * bs, dtype=torch.float16, device=DEV)
    st = torch.cuda.current_stream()
    for rounds in range(3):
        for i in range(9):
            assert L.ggufb200_dequant(int(qt), devs[i % 3].data_ptr(), n_blocks, out.data_ptr(), 0, lib.DEQUANT_SRC_STABLE, st.cuda_stream) == 0
        torch.cuda.synchronize()
        assert np.array_equal(torch_bits(out), oracle.dequant(raws[2], int(qt), oracle.DT_F16, oracle.DT_F16))


def test_src_stable_after_a_torch_kernel_that_writes_the_packed_bytes(pkg):
    """dequantize() passes SRC_STABLE by default (dequant.py): the packed bytes come from loads, copies or ordinary torch
    kernels, which do not signal programmatic completion early, so they are complete before our kernel starts.  Here a torch
    kernel rewrites a 37 MB packed tensor IN PLACE and the dequant follows immediately on the same stream, 24 times with
    different contents: the result must always be the dequant of the new bytes."""
    qt = Q.Q4_K
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = 21504 * 3072 // bs
    seed = oracle.random_blocks(int(qt), 1 << 12, seed=21, scale=0.02)
    base = torch.from_numpy(seed).to(DEV).repeat(n_blocks // (1 << 12), 1).contiguous()
    # flipping quant bytes only (offsets >= 16) keeps every fp16 field finite
    mask = torch.zeros(ts, dtype=torch.uint8, device=DEV)
    work = base.clone()
    torch.cuda.synchronize()
    for i in range(24):
        mask.zero_()
        mask[16 + (i * 5) % 128] = 1 + (i % 255)
        torch.bitwise_xor(base, mask, out=work)                          # the writer: an ordinary torch kernel over 37 MB
        got = pkg.dequant.dequantize(work, qt, (n_blocks * bs,), src_stable=True)
        torch.cuda.synchronize()
        want = pkg.dequant.dequantize(work, qt, (n_blocks * bs,), src_stable=False)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), f"iteration {i}: stale packed bytes were read"
