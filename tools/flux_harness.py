"""Flux.1-dev-SHAPED DiT for the denoise-step benchmark (BASELINE.json configs[2] / configs[4]).

Benchmark scaffolding, not part of the product package: ComfyUI (which owns the real model code) is not available
offline, so this file rebuilds the Flux.1-dev block structure -- 19 double-stream + 38 single-stream blocks, hidden
3072, 24 heads x 128, mlp 12288, every nn.Linear created through an injected `operations` namespace exactly the way
ComfyUI injects `custom_operations` -- with random-init weights packed as GGUF Q4_K (the 304 block Linears, what a
Q4_K_S file contains) and BF16 (in/out/embedding Linears), F32 biases.  One forward = one denoise step.

Two interchangeable `operations`:
    ours        comfyui-gguf_b200 GGMLOps             (fused / tensor-core kernels of this repo)
    reference   RefChainOps: the reference's torch path restated op-for-op (oracle/torch_chain.py)
Both read the SAME packed weight tensors, so outputs can be compared and timed A/B in one process.
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F
import gguf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Q = gguf.GGMLQuantizationType


# ------------------------------------------------------------------ synthetic packed weights, generated on the GPU
_F16_FIELDS = {Q.Q4_K: (0, 2), Q.Q5_K: (0, 2), Q.Q6_K: (208,), Q.Q8_0: (0,), Q.Q4_0: (0,), Q.Q2_K: (80, 82), Q.Q3_K: (108,)}


def random_packed(qtype, N, K, device, gen, scale=2e-4):
    """Random payload bytes with finite, small fp16 scale fields (SURVEY.md 8d recipe), built with torch on `device`."""
    bs, ts = gguf.GGML_QUANT_SIZES[qtype]
    if qtype == Q.BF16:
        w = torch.randn(N, K, device=device, generator=gen) * 0.02
        return w.to(torch.bfloat16).view(torch.uint8).reshape(N, K * 2)
    n_blocks = N * K // bs
    raw = torch.randint(0, 256, (n_blocks, ts), dtype=torch.uint8, device=device, generator=gen)
    s = scale if qtype not in (Q.Q8_0, Q.Q4_0) else scale * 20
    for off in _F16_FIELDS[qtype]:
        f = (torch.randn(n_blocks, device=device, generator=gen) * s).to(torch.float16)
        raw[:, off:off + 2] = f.view(torch.uint8).reshape(n_blocks, 2)
    return raw.reshape(N, K // bs * ts)


# ------------------------------------------------------------------ the reference's execution path as an `operations` namespace
class RefChainOps:
    """`custom_operations` whose Linear runs the reference's unfused torch chain (dequant.py ops -> F.linear)."""

    class Linear(nn.Module):
        def __init__(self, in_features, out_features, bias=True, device=None, dtype=None):
            super().__init__()
            self.in_features, self.out_features = in_features, out_features
            self.weight = None
            self.bias = None
            self.dequant_dtype = None

        def forward(self, x):
            from oracle import torch_chain
            w = self.weight
            qt = getattr(w, "tensor_type", None)
            data = w.as_subclass(torch.Tensor) if type(w) is not torch.Tensor else w
            b = None if self.bias is None else (self.bias.as_subclass(torch.Tensor) if type(self.bias) is not torch.Tensor else self.bias)
            return torch_chain.linear(x, data, int(qt), tuple(w.tensor_shape), b, self.dequant_dtype)


# ------------------------------------------------------------------ model
def rope_table(L, dim, device, theta=10000.0):
    pos = torch.arange(L, device=device, dtype=torch.float32)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
    return torch.polar(torch.ones(L, dim // 2, device=device), torch.outer(pos, freqs))[None, None]   # complex64 [1,1,L,dim/2]


def apply_rope(x, fc):
    xc = torch.view_as_complex(x.float().unflatten(-1, (-1, 2)))
    return torch.view_as_real(xc * fc).flatten(-2).to(x.dtype)


def rms(x, scale, eps=1e-6):
    return F.rms_norm(x, (x.shape[-1],), scale.to(x.dtype), eps)


def modulate(x, shift, scale):
    return (1 + scale) * F.layer_norm(x, x.shape[-1:], eps=1e-6) + shift


def attention(q, k, v, cs):
    q, k = apply_rope(q, cs), apply_rope(k, cs)
    o = F.scaled_dot_product_attention(q, k, v)
    return o.transpose(1, 2).flatten(2)


class DoubleBlock(nn.Module):
    def __init__(self, ops, h, heads, mlp):
        super().__init__()
        self.heads = heads
        for s in ("img", "txt"):
            setattr(self, f"{s}_mod", ops.Linear(h, 6 * h))
            setattr(self, f"{s}_qkv", ops.Linear(h, 3 * h))
            setattr(self, f"{s}_proj", ops.Linear(h, h))
            setattr(self, f"{s}_mlp0", ops.Linear(h, mlp))
            setattr(self, f"{s}_mlp2", ops.Linear(mlp, h))
            self.register_buffer(f"{s}_qs", torch.ones(h // heads), persistent=False)
            self.register_buffer(f"{s}_ks", torch.ones(h // heads), persistent=False)

    def _qkv(self, s, x, mod):
        B, L, _ = x.shape
        qkv = getattr(self, f"{s}_qkv")(modulate(x, mod[0], mod[1]))
        q, k, v = qkv.view(B, L, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        return rms(q, getattr(self, f"{s}_qs")), rms(k, getattr(self, f"{s}_ks")), v

    def forward(self, img, txt, vec, cs):
        im = self.img_mod(F.silu(vec))[:, None].chunk(6, dim=-1)
        tm = self.txt_mod(F.silu(vec))[:, None].chunk(6, dim=-1)
        iq, ik, iv = self._qkv("img", img, im)
        tq, tk, tv = self._qkv("txt", txt, tm)
        a = attention(torch.cat((tq, iq), 2), torch.cat((tk, ik), 2), torch.cat((tv, iv), 2), cs)
        ta, ia = a[:, :txt.shape[1]], a[:, txt.shape[1]:]
        img = img + im[2] * self.img_proj(ia)
        img = img + im[5] * self.img_mlp2(F.gelu(self.img_mlp0(modulate(img, im[3], im[4])), approximate="tanh"))
        txt = txt + tm[2] * self.txt_proj(ta)
        txt = txt + tm[5] * self.txt_mlp2(F.gelu(self.txt_mlp0(modulate(txt, tm[3], tm[4])), approximate="tanh"))
        return img, txt


class SingleBlock(nn.Module):
    def __init__(self, ops, h, heads, mlp):
        super().__init__()
        self.heads, self.h, self.mlp = heads, h, mlp
        self.modulation = ops.Linear(h, 3 * h)
        self.linear1 = ops.Linear(h, 3 * h + mlp)
        self.linear2 = ops.Linear(h + mlp, h)
        self.register_buffer("qs", torch.ones(h // heads), persistent=False)
        self.register_buffer("ks", torch.ones(h // heads), persistent=False)

    def forward(self, x, vec, cs):
        shift, scale, gate = self.modulation(F.silu(vec))[:, None].chunk(3, dim=-1)
        B, L, _ = x.shape
        qkv, m = torch.split(self.linear1(modulate(x, shift, scale)), [3 * self.h, self.mlp], dim=-1)
        q, k, v = qkv.view(B, L, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        a = attention(rms(q, self.qs), rms(k, self.ks), v, cs)
        return x + gate * self.linear2(torch.cat((a, F.gelu(m, approximate="tanh")), 2))


class FluxShapeDiT(nn.Module):
    def __init__(self, ops, hidden=3072, heads=24, mlp_ratio=4, depth=19, depth_single=38, in_ch=64, ctx=4096, vec=768):
        super().__init__()
        mlp = hidden * mlp_ratio
        self.hidden, self.heads = hidden, heads
        self.img_in = ops.Linear(in_ch, hidden)
        self.txt_in = ops.Linear(ctx, hidden)
        self.time_in0, self.time_in2 = ops.Linear(256, hidden), ops.Linear(hidden, hidden)
        self.vector_in0, self.vector_in2 = ops.Linear(vec, hidden), ops.Linear(hidden, hidden)
        self.guidance_in0, self.guidance_in2 = ops.Linear(256, hidden), ops.Linear(hidden, hidden)
        self.double_blocks = nn.ModuleList([DoubleBlock(ops, hidden, heads, mlp) for _ in range(depth)])
        self.single_blocks = nn.ModuleList([SingleBlock(ops, hidden, heads, mlp) for _ in range(depth_single)])
        self.final_mod = ops.Linear(hidden, 2 * hidden)
        self.final = ops.Linear(hidden, in_ch)

    @staticmethod
    def _temb(t, dim=256):
        half = dim // 2
        f = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
        a = t[:, None].float() * 1000.0 * f[None]
        return torch.cat((torch.cos(a), torch.sin(a)), -1)

    def forward(self, img, txt, t, y, guidance):
        dt = img.dtype
        vec = self.time_in2(F.silu(self.time_in0(self._temb(t).to(dt))))
        vec = vec + self.guidance_in2(F.silu(self.guidance_in0(self._temb(guidance).to(dt))))
        vec = vec + self.vector_in2(F.silu(self.vector_in0(y)))
        img, txt = self.img_in(img), self.txt_in(txt)
        cs = rope_table(txt.shape[1] + img.shape[1], self.hidden // self.heads, img.device)
        for blk in self.double_blocks:
            img, txt = blk(img, txt, vec, cs)
        x = torch.cat((txt, img), 1)
        for blk in self.single_blocks:
            x = blk(x, vec, cs)
        x = x[:, txt.shape[1]:]
        shift, scale = self.final_mod(F.silu(vec))[:, None].chunk(2, dim=-1)
        return self.final(modulate(x, shift, scale))


BLOCK_LINEARS = ("_mod", "_qkv", "_proj", "_mlp0", "_mlp2", "modulation", "linear1", "linear2")


def build_state_dict(model, GGMLTensor, device, seed=0, block_qtype=Q.Q4_K, other_qtype=Q.BF16):
    """Random packed weights for every Linear of `model`: block Linears `block_qtype`, the rest `other_qtype`; F32 biases."""
    gen = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, mod in model.named_modules():
        if not hasattr(mod, "in_features"):
            continue
        N, K = mod.out_features, mod.in_features
        is_block = any(name.endswith(s) or s in name.split(".")[-1] for s in BLOCK_LINEARS) and ("double_blocks" in name or "single_blocks" in name)
        qt = block_qtype if is_block else other_qtype
        bs, _ = gguf.GGML_QUANT_SIZES[qt]
        if K % bs != 0:
            qt = Q.BF16
        raw = random_packed(qt, N, K, device, gen)
        sd[f"{name}.weight"] = GGMLTensor(raw, tensor_type=qt, tensor_shape=torch.Size((N, K)))
        b = torch.randn(N, device=device, generator=gen) * 0.02
        sd[f"{name}.bias"] = GGMLTensor(b, tensor_type=Q.F32, tensor_shape=torch.Size((N,)))
    return sd


def load_shared(model, sd):
    """Attach the SAME tensor objects to a model (both arms read identical packed bytes)."""
    for name, mod in model.named_modules():
        if hasattr(mod, "in_features"):
            mod.weight = nn.Parameter(sd[f"{name}.weight"], requires_grad=False)
            mod.bias = nn.Parameter(sd[f"{name}.bias"], requires_grad=False)
    return model


def make_inputs(device, dtype, batch=1, img_tokens=4096, txt_tokens=512, seed=1):
    g = torch.Generator(device=device).manual_seed(seed)
    return dict(
        img=torch.randn(batch, img_tokens, 64, device=device, generator=g).to(dtype),
        txt=torch.randn(batch, txt_tokens, 4096, device=device, generator=g).to(dtype),
        t=torch.full((batch,), 0.5, device=device),
        y=torch.randn(batch, 768, device=device, generator=g).to(dtype),
        guidance=torch.full((batch,), 3.5, device=device),
    )


def linear_flops(model, img_tokens, txt_tokens, batch=1):
    """2*M*N*K summed over every Linear call of one forward."""
    total = 0
    for name, mod in model.named_modules():
        if not hasattr(mod, "in_features"):
            continue
        leaf = name.split(".")[-1]
        if leaf.startswith("img_") and not leaf.endswith("_mod") and leaf != "img_in":
            M = img_tokens
        elif leaf.startswith("txt_") and not leaf.endswith("_mod") and leaf != "txt_in":
            M = txt_tokens
        elif leaf in ("linear1", "linear2"):
            M = img_tokens + txt_tokens
        elif leaf == "img_in" or leaf == "final":
            M = img_tokens
        elif leaf == "txt_in":
            M = txt_tokens
        else:
            M = 1
        total += 2 * batch * M * mod.in_features * mod.out_features
    return total


# ------------------------------------------------------------------ per-Linear parity at model scale
class LinearParity:
    """Forward hooks on every quantised Linear of `model`: the layer's output (whatever route AUTO picked for it) against the
    reference arithmetic ON THE SAME INPUT -- fp32-accumulated x @ W^T + bias with W = the bit-exact standalone dequant
    (dequant.py float sequence) rounded to the activation dtype, bias rounded to the activation dtype (ops.py:193-211,
    242-244), result rounded to the activation dtype.  Isolates the per-layer error from the drift a deep network
    accumulates: `records` = (name, qtype, M, N, K, relative Frobenius error)."""

    def __init__(self, model, dequant_mod, every=1):
        self.model, self.dq, self.every = model, dequant_mod, every
        self.records, self.handles = [], []

    def __enter__(self):
        idx = 0
        for name, mod in self.model.named_modules():
            w = getattr(mod, "weight", None)
            if not hasattr(mod, "in_features") or w is None or getattr(w, "tensor_type", None) is None:
                continue
            idx += 1
            if idx % self.every:
                continue
            self.handles.append(mod.register_forward_hook(self._hook(name)))
        return self

    def _hook(self, name):
        def fn(mod, args, out):
            x = args[0]
            if not x.is_cuda or x.dtype not in (torch.float16, torch.bfloat16):
                return
            w = mod.weight
            K = x.shape[-1]
            x2 = x.reshape(-1, K)
            W = self.dq.dequantize_tensor(w, x.dtype, getattr(mod, "dequant_dtype", None)).as_subclass(torch.Tensor)
            ref = x2.float() @ W.float().t()
            b = getattr(mod, "bias", None)
            if b is not None:
                ref = ref + b.as_subclass(torch.Tensor).to(x.device).to(x.dtype).float()
            ref = ref.to(x.dtype).float()
            y = out.reshape(-1, out.shape[-1]).float()
            rel = float(((y - ref).norm() / ref.norm().clamp_min(1e-30)).item())
            self.records.append((name, w.tensor_type.name, x2.shape[0], W.shape[0], K, rel))
            del W, ref
        return fn

    def __exit__(self, *exc):
        for h in self.handles:
            h.remove()

    def summary(self):
        import collections
        by = collections.defaultdict(list)
        for _n, qt, M, _N, _K, rel in self.records:
            by[(qt, "M<=8" if M <= 8 else ("M<=1024" if M <= 1024 else "M>1024"))].append(rel)
        rows = {f"{qt} {cls}": {"layers": len(v), "max": max(v), "median": sorted(v)[len(v) // 2]} for (qt, cls), v in sorted(by.items())}
        worst = max(self.records, key=lambda r: r[-1]) if self.records else None
        return {"per_class": rows, "worst": worst, "layers": len(self.records)}
