"""GGUF file -> state dict of GGMLTensor (drop-in for the reference's loader.py, hot-path part only).

Mirrored (reference file:line): gguf_sd_loader loader.py:51-141 (mmap read, prefix strip, architecture check,
`comfy.gguf.orig_shape.*` metadata loader.py:16-24, F32/F16 reshaped views vs raw uint8 payloads loader.py:118-120,
1-D BF16 -> F32 loader.py:122-124, qtype histogram log loader.py:130-131, largest-weight mark loader.py:133-137)
and the T5 / llama key remapping of gguf_clip_loader loader.py:377-406.

Out of scope here: the sd.cpp "compat" architecture sniffing (needs
tools/convert.py), mmproj / vision towers and tokenizer reconstruction -- those raise NotImplementedError.
"""
from __future__ import annotations

import logging
import warnings

import gguf
import torch

from .dequant import dequantize_tensor, is_quantized
from .ops import GGMLTensor

IMG_ARCH_LIST = {"flux", "sd1", "sdxl", "sd3", "aura", "hidream", "cosmos", "ltxv", "hyvid", "wan", "lumina2", "qwen_image"}
TXT_ARCH_LIST = {"t5", "t5encoder", "llama", "qwen2vl", "qwen3", "qwen3vl"}
VIS_TYPE_LIST = {"clip-vision", "mmproj"}
_Q = gguf.GGMLQuantizationType


def _string_field(reader, name):
    field = reader.get_field(name)
    if field is None:
        return None
    if len(field.types) != 1 or field.types[0] != gguf.GGUFValueType.STRING:
        raise TypeError(f"Bad type for GGUF {name} key: expected string, got {field.types!r}")
    return str(field.parts[field.data[-1]], encoding="utf-8")


def get_orig_shape(reader, tensor_name):
    """Logical shape stored by the converter for tensors it had to reshape (loader.py:16-24)."""
    key = f"comfy.gguf.orig_shape.{tensor_name}"
    field = reader.get_field(key)
    if field is None:
        return None
    if len(field.types) != 2 or field.types[0] != gguf.GGUFValueType.ARRAY or field.types[1] != gguf.GGUFValueType.INT32:
        raise TypeError(f"Bad original shape metadata for {key}: Expected ARRAY of INT32, got {field.types}")
    return torch.Size(int(field.parts[i][0]) for i in field.data)


def _check_arch(arch, kind, is_text_model, path):
    if arch in (None, "pig", "cow"):
        if is_text_model:
            raise ValueError(f"This gguf file is incompatible with llama.cpp!\nConsider using safetensors or a compatible gguf file\n({path})")
        raise NotImplementedError("stable-diffusion.cpp style files without general.architecture need the converter's "
                                  "architecture sniffing, which is outside this package's scope")
    if is_text_model:
        if arch not in TXT_ARCH_LIST and kind not in VIS_TYPE_LIST:
            raise ValueError(f"Unexpected text model architecture type in GGUF file: {arch!r}")
    elif arch not in IMG_ARCH_LIST:
        raise ValueError(f"Unexpected architecture type in GGUF file: {arch!r}")


def gguf_sd_loader(path, handle_prefix="model.diffusion_model.", return_arch=False, is_text_model=False):
    """Read a GGUF file into {key: GGMLTensor}; quantised payloads stay packed uint8 views of the mmap."""
    reader = gguf.GGUFReader(path)

    entries = [(t.name, t) for t in reader.tensors]
    if handle_prefix is not None and any(name.startswith(handle_prefix) for name, _ in entries):
        cut = len(handle_prefix)
        entries = [(name[cut:], t) for name, t in entries if name.startswith(handle_prefix)]

    arch = _string_field(reader, "general.architecture")
    _check_arch(arch, _string_field(reader, "general.type"), is_text_model, path)

    state_dict, histogram = {}, {}
    for key, t in entries:
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
            payload = torch.from_numpy(t.data)  # zero-copy view of the mmap
        shape = get_orig_shape(reader, t.name)
        if shape is None:
            shape = torch.Size(int(v) for v in reversed(t.shape))
        if t.tensor_type in (_Q.F32, _Q.F16):
            payload = payload.view(*shape)
        item = GGMLTensor(payload, tensor_type=t.tensor_type, tensor_shape=shape)
        if len(shape) <= 1 and t.tensor_type == _Q.BF16:
            # 1-D tensors are never meant to be quantised: plain widening cast bf16 -> fp32 at load time
            item = payload.view(torch.bfloat16).to(torch.float32).reshape(shape)
        state_dict[key] = item
        tname = getattr(t.tensor_type, "name", repr(t.tensor_type))
        histogram[tname] = histogram.get(tname, 0) + 1

    logging.info("gguf qtypes: " + ", ".join(f"{k} ({v})" for k, v in histogram.items()))

    quantised = [k for k, v in state_dict.items() if is_quantized(v)]
    if quantised:
        biggest = max(quantised, key=lambda k: state_dict[k].numel())
        state_dict[biggest].is_largest_weight = True   # read by GGMLLayer for VRAM estimation

    return (state_dict, arch) if return_arch else state_dict


# llama.cpp tensor names -> original checkpoint names (order matters: longer patterns first where they overlap)
T5_SD_MAP = (
    ("enc.", "encoder."), (".blk.", ".block."), ("token_embd", "shared"), ("output_norm", "final_layer_norm"),
    ("attn_q", "layer.0.SelfAttention.q"), ("attn_k", "layer.0.SelfAttention.k"), ("attn_v", "layer.0.SelfAttention.v"),
    ("attn_o", "layer.0.SelfAttention.o"), ("attn_norm", "layer.0.layer_norm"),
    ("attn_rel_b", "layer.0.SelfAttention.relative_attention_bias"),
    ("ffn_up", "layer.1.DenseReluDense.wi_1"), ("ffn_down", "layer.1.DenseReluDense.wo"),
    ("ffn_gate", "layer.1.DenseReluDense.wi_0"), ("ffn_norm", "layer.1.layer_norm"),
)
LLAMA_SD_MAP = (
    ("blk.", "model.layers."), ("attn_norm", "input_layernorm"), ("attn_q_norm.", "self_attn.q_norm."),
    ("attn_k_norm.", "self_attn.k_norm."), ("attn_v_norm.", "self_attn.v_norm."), ("attn_q", "self_attn.q_proj"),
    ("attn_k", "self_attn.k_proj"), ("attn_v", "self_attn.v_proj"), ("attn_output", "self_attn.o_proj"),
    ("ffn_up", "mlp.up_proj"), ("ffn_down", "mlp.down_proj"), ("ffn_gate", "mlp.gate_proj"),
    ("ffn_norm", "post_attention_layernorm"), ("token_embd", "model.embed_tokens"), ("output_norm", "model.norm"),
    ("output.weight", "lm_head.weight"),
)


def sd_map_replace(raw_sd, key_map):
    pairs = key_map.items() if isinstance(key_map, dict) else key_map
    pairs = list(pairs)
    out = {}
    for key, value in raw_sd.items():
        for old, new in pairs:
            key = key.replace(old, new)
        out[key] = value
    return out


def llama_permute(raw_sd, n_head, n_head_kv):
    """Undo llama.cpp's rotary-friendly q/k row permutation (loader.py:205-216); acts on the packed rows."""
    def unpermute(x, heads):
        return x.reshape(heads, x.shape[0] // heads // 2, 2, *x.shape[1:]).swapaxes(1, 2).reshape(x.shape)
    for key, value in raw_sd.items():
        if key.endswith(("q_proj.weight", "q_proj.bias")):
            value.data = unpermute(value.data, n_head)
        elif key.endswith(("k_proj.weight", "k_proj.bias")):
            value.data = unpermute(value.data, n_head_kv)
    return raw_sd


def gguf_clip_loader(path):
    """Text-encoder GGUF -> state dict with original key names (loader.py:377-406, T5 and llama-family parts)."""
    sd, arch = gguf_sd_loader(path, return_arch=True, is_text_model=True)
    temb = "token_embd.weight"
    if arch in {"t5", "t5encoder"}:
        if temb in sd and tuple(sd[temb].shape) == (256384, 4096):
            raise NotImplementedError("Comfy-Org T5 tokenizer reconstruction is outside this package's scope")
        return sd_map_replace(sd, T5_SD_MAP)
    if arch in {"llama", "qwen2vl", "qwen3", "qwen3vl"}:
        if arch == "qwen2vl":
            raise NotImplementedError("mmproj / vision tower loading is outside this package's scope")
        if temb in sd and sd[temb].shape[0] >= (64 * 1024):
            if arch == "llama" and tuple(sd[temb].shape) == (131072, 5120):
                raise NotImplementedError("tekken tokenizer reconstruction is outside this package's scope")
            # the reference pre-dequantises huge embedding tables to dodge its whole-table dequant per call
            # (loader.py:391-397); the row-gather kernel makes that unnecessary, but the host model may index
            # the table directly, so keep the reference behaviour
            logging.warning(f"Dequantizing {temb} to prevent runtime OOM.")
            sd[temb] = dequantize_tensor(sd[temb], dtype=torch.float16)
        sd = sd_map_replace(sd, LLAMA_SD_MAP)
        if arch == "llama":
            sd = llama_permute(sd, 32, 8)
        return sd
    return sd
