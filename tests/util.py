"""Shared helpers for the parity tests."""
import numpy as np
import torch
import gguf

Q = gguf.GGMLQuantizationType
ALL_QTYPES = [Q.Q4_0, Q.Q4_1, Q.Q5_0, Q.Q5_1, Q.Q8_0, Q.Q2_K, Q.Q3_K, Q.Q4_K, Q.Q5_K, Q.Q6_K, Q.IQ4_NL, Q.IQ4_XS, Q.BF16]
TORCH_DT = {0: torch.float16, 1: torch.bfloat16, 2: torch.float32}
COMBOS = [(0, 0), (0, 1), (0, 2), (1, 1), (2, 2), (2, 0)]   # (math, out) pairs stored in the golden files


def canon_nan(bits: np.ndarray, out_dtype: int) -> np.ndarray:
    """Map every NaN encoding to one canonical pattern (the payload of a NaN is not part of the contract)."""
    bits = bits.copy()
    if out_dtype == 2:
        u = bits.view(np.uint32)
        u[(u & 0x7FFFFFFF) > 0x7F800000] = 0x7FC00000
        return u
    if out_dtype == 0:
        bits[(bits & 0x7FFF) > 0x7C00] = 0x7E00
    else:
        bits[(bits & 0x7FFF) > 0x7F80] = 0x7FC0
    return bits


def torch_bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous().reshape(-1)
    if t.dtype == torch.float32:
        return t.numpy().view(np.uint32).copy()
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def bits_to_f32(bits: np.ndarray, dtype_code: int) -> np.ndarray:
    if dtype_code == 2:
        return bits.view(np.float32) if bits.dtype != np.float32 else bits
    if dtype_code == 0:
        return bits.view(np.float16).astype(np.float32)
    return (bits.astype(np.uint32) << 16).view(np.float32)


def rel_fro(a: np.ndarray, b: np.ndarray) -> float:
    a = a.astype(np.float64).reshape(-1)
    b = b.astype(np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
