"""CPU tests: the oracle (oracle/gguf_oracle.c) is pinned against the reference-generated golden vectors and gguf-py."""
import os

import numpy as np
import pytest
import gguf

import oracle
from util import ALL_QTYPES, COMBOS, Q, bits_to_f32, canon_nan, rel_fro


@pytest.mark.parametrize("qt", ALL_QTYPES, ids=lambda q: q.name)
def test_oracle_matches_reference_golden_bit_exact(qt, golden_dir):
    g = np.load(os.path.join(golden_dir, f"dequant_{qt.name}.npz"))
    packed = g["packed"]
    for math, od in COMBOS:
        want = canon_nan(g[f"out_m{math}_o{od}"], od)
        got = oracle.dequant(packed, int(qt), od, math)
        got = canon_nan(got.view(np.uint32) if od == 2 else got, od)
        assert np.array_equal(got, want), f"{qt.name} math={math} out={od}: {np.count_nonzero(got != want)} mismatches"


@pytest.mark.parametrize("qt", ALL_QTYPES, ids=lambda q: q.name)
def test_oracle_matches_gguf_py_numpy(qt):
    """gguf.quants.dequantize is the reference's own fallback path (dequant.py:24-28): fp32 math, fp32 out."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    n_blocks = 257 if bs > 1 else 5003
    raw = oracle.random_blocks(int(qt), n_blocks, seed=11)
    want = gguf.quants.dequantize(raw.reshape(-1) if bs == 1 else raw, qt).reshape(-1)
    got = oracle.dequant(raw, int(qt), oracle.DT_F32, oracle.DT_F32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("qt", [q for q in ALL_QTYPES if q != Q.BF16], ids=lambda q: q.name)
def test_oracle_integer_unpack_consistent_with_fp32_dequant(qt):
    """out = f(d, d2, q, sc, mn) in exact fp32 arithmetic: the integer unpack must explain the float result."""
    bs, ts = gguf.GGML_QUANT_SIZES[qt]
    raw = oracle.random_blocks(int(qt), 64, seed=5)
    q, sc, mn = oracle.unpack_int(raw, int(qt))
    assert q.shape == (64 * bs,)
    out = oracle.dequant(raw, int(qt), oracle.DT_F32, oracle.DT_F32).reshape(64, bs)
    offs = oracle._F16_FIELDS[int(qt)]
    d = raw[:, offs[0]:offs[0] + 2].copy().view(np.float16).astype(np.float32)
    d2 = raw[:, offs[1]:offs[1] + 2].copy().view(np.float16).astype(np.float32) if len(offs) > 1 else np.zeros_like(d)
    qf, scf, mnf = (a.reshape(64, bs).astype(np.float32) for a in (q, sc, mn))
    if qt in (Q.Q4_1, Q.Q5_1):
        want = d * qf + d2
    elif qt in (Q.Q2_K, Q.Q4_K, Q.Q5_K):
        want = (d * scf) * qf - (d2 * mnf)
    else:
        want = (d * scf) * qf
    assert np.array_equal(want.view(np.uint32), out.view(np.uint32))


def test_oracle_empty_and_bad_type():
    assert oracle.dequant(np.zeros(0, np.uint8), int(Q.Q4_K)).size == 0
    with pytest.raises(ValueError):
        oracle.type_info(999)


@pytest.mark.parametrize("name", ["Q4_K", "Q8_0", "Q5_K", "Q6_K", "Q4_0", "BF16"])
@pytest.mark.parametrize("act,code", [("bf16", 1), ("f16", 0), ("f32", 2)])
def test_oracle_linear_matches_reference_ops_golden(name, act, code, golden_dir):
    """Golden y comes from the unmodified reference GGMLOps.Linear on CPU.  Tolerance: 1e-3 relative (Frobenius),
    the Linear contract of BASELINE.json; bit equality is not expected because accumulation order differs."""
    g = np.load(os.path.join(golden_dir, f"linear_{name}_{act}.npz"))
    N, K, M = int(g["N"]), int(g["K"]), int(g["M"])
    bias32 = g["bias"]
    # the reference casts the F32 bias to the activation dtype first
    if code == 2:
        bias_bits = bias32
    elif code == 0:
        bias_bits = bias32.astype(np.float16).view(np.uint16)
    else:
        import torch
        bias_bits = torch.from_numpy(bias32).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    y = oracle.linear(g["packed"], int(g["qtype"]), N, K, g["x"], code, oracle.DT_F16, bias_bits)
    got = bits_to_f32(y.reshape(-1), code)
    want = bits_to_f32(g["y"], code)
    assert rel_fro(got, want) <= 1e-3
