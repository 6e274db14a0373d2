"""CPU oracle for the GGUF dequant + Linear path -- TEST INFRASTRUCTURE ONLY.

Loads ``oracle/libgguf_oracle.so`` (built from ``gguf_oracle.c``; see that file's
header for what it restates and how parity is pinned) and exposes numpy-level
helpers.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgguf_oracle.so")
_lib = None

DT_F16, DT_BF16, DT_F32 = 0, 1, 2
_NP_OUT = {DT_F16: np.uint16, DT_BF16: np.uint16, DT_F32: np.float32}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (used by __graft_entry__.build())."""
    src = os.path.join(_HERE, "gguf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgguf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.ggor_type_info.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.ggor_dequant.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.ggor_unpack_int.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ggor_linear.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.ggor_num_threads.restype = ctypes.c_int
        L.ggor_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def type_info(qtype: int) -> tuple[int, int]:
    bs, ts = ctypes.c_int(), ctypes.c_int()
    if lib().ggor_type_info(int(qtype), ctypes.byref(bs), ctypes.byref(ts)) != 0:
        raise ValueError(f"oracle: unknown ggml type {qtype}")
    return bs.value, ts.value


def dequant(packed: np.ndarray, qtype: int, out_dtype: int = DT_F16, math_dtype: int = DT_F16, out: np.ndarray | None = None) -> np.ndarray:
    """Dequantise a flat uint8 block stream.  fp16/bf16 results come back as raw uint16 bit patterns.
    `out`: optional preallocated flat result array (bench.py's reference arm reuses its output buffers like the GPU arm does)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1)
    bs, ts = type_info(qtype)
    assert packed.size % ts == 0, (packed.size, ts)
    n_blocks = packed.size // ts
    if out is None:
        out = np.empty(n_blocks * bs, dtype=_NP_OUT[out_dtype])
    assert out.dtype == _NP_OUT[out_dtype] and out.size == n_blocks * bs and out.flags.c_contiguous
    rc = lib().ggor_dequant(int(qtype), packed.ctypes.data, n_blocks, out.ctypes.data, out_dtype, math_dtype)
    if rc != 0:
        raise RuntimeError(f"oracle dequant failed rc={rc}")
    return out


def unpack_int(packed: np.ndarray, qtype: int):
    """Integer unpack (q, sc, mn) per element as int16 arrays."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1)
    bs, ts = type_info(qtype)
    n_blocks = packed.size // ts
    q = np.empty(n_blocks * bs, dtype=np.int16)
    sc = np.empty_like(q)
    mn = np.empty_like(q)
    rc = lib().ggor_unpack_int(int(qtype), packed.ctypes.data, n_blocks, q.ctypes.data, sc.ctypes.data, mn.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle unpack failed rc={rc}")
    return q, sc, mn


def linear(packed: np.ndarray, qtype: int, N: int, K: int, x_bits: np.ndarray, act_dtype: int,
           math_dtype: int = DT_F16, bias_bits: np.ndarray | None = None) -> np.ndarray:
    """y = x @ dequant(W).T + bias with x/bias/y given as act-dtype bit patterns (uint16) or float32."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1)
    x_bits = np.ascontiguousarray(x_bits)
    M = x_bits.size // K
    y = np.empty(M * N, dtype=_NP_OUT[act_dtype])
    bptr = None
    if bias_bits is not None:
        bias_bits = np.ascontiguousarray(bias_bits)
        bptr = bias_bits.ctypes.data
    rc = lib().ggor_linear(int(qtype), packed.ctypes.data, N, K, x_bits.ctypes.data, M, act_dtype, math_dtype, bptr, y.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle linear failed rc={rc}")
    return y.reshape(M, N)


def num_threads() -> int:
    return lib().ggor_num_threads()


def set_num_threads(n: int) -> None:
    lib().ggor_set_num_threads(int(n))


# ---- helpers shared by tests / bench: seeded synthetic packed tensors ----
_F16_FIELDS = {  # byte offsets of every fp16 header field per block
    2: (0,), 3: (0, 2), 6: (0,), 7: (0, 2), 8: (0,), 20: (0,),
    10: (80, 82), 11: (108,), 12: (0, 2), 13: (0, 2), 14: (208,), 23: (0,),
}


def random_blocks(qtype: int, n_blocks: int, seed: int = 0, scale: float = 0.01) -> np.ndarray:
    """Uniform random payload bytes with every fp16 header field overwritten by a finite fp16 N(0, scale^2)."""
    bs, ts = type_info(qtype)
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n_blocks, ts), dtype=np.uint8)
    if qtype == 30:
        vals = rng.normal(0.0, 0.02, size=n_blocks).astype(np.float32)
        return (vals.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8).reshape(n_blocks, 2)
    for off in _F16_FIELDS[int(qtype)]:
        f = rng.normal(0.0, scale, size=n_blocks).astype(np.float16)
        raw[:, off:off + 2] = f.view(np.uint8).reshape(n_blocks, 2)
    return raw
