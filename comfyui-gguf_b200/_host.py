"""Resolution of the ComfyUI host interfaces this package plugs into.

Inside ComfyUI the real `comfy.ops`, `comfy.lora` and `comfy.model_management` are used.  Outside of it
(benchmarks, unit tests without the fake package, the GPU box) a minimal stand-in with the same call
signatures is provided so the op classes stay importable and runnable.  The stand-in implements only
what the GGUF op layer calls (reference ops.py:186-210, 227-271).
"""
from __future__ import annotations

import types

import torch

try:  # pragma: no cover - exercised only inside ComfyUI / with tests/fake_comfy on sys.path
    import comfy.ops as comfy_ops
    import comfy.lora as comfy_lora
    import comfy.model_management as comfy_mm
    HAVE_COMFY = True
except ImportError:
    HAVE_COMFY = False

    def _cast_to(weight, dtype=None, device=None, non_blocking=False, copy=False):
        if weight is None:
            return None
        same_dev = device is None or weight.device == torch.device(device)
        if same_dev and not copy and (dtype is None or weight.dtype == dtype):
            return weight
        return weight.to(device=device, dtype=dtype, non_blocking=non_blocking, copy=copy)

    def _cast_bias_weight(s, input=None, dtype=None, device=None, bias_dtype=None):
        if input is not None:
            dtype = dtype or input.dtype
            bias_dtype = bias_dtype or dtype
            device = device or input.device
        bias = _cast_to(getattr(s, "bias", None), bias_dtype, device)
        return _cast_to(s.weight, dtype, device), bias

    class _CastOp:
        comfy_cast_weights = False
        weight_function = []
        bias_function = []

        def reset_parameters(self):
            return None

        def forward(self, *args, **kwargs):
            if self.comfy_cast_weights or self.weight_function or self.bias_function:
                return self.forward_comfy_cast_weights(*args, **kwargs)
            return super().forward(*args, **kwargs)

    def _family(cast):
        ns = {}

        class Linear(_CastOp, torch.nn.Linear):
            comfy_cast_weights = cast

            def forward_comfy_cast_weights(self, input):
                w, b = _cast_bias_weight(self, input)
                return torch.nn.functional.linear(input, w, b)

        class Conv2d(_CastOp, torch.nn.Conv2d):
            comfy_cast_weights = cast

            def forward_comfy_cast_weights(self, input):
                w, b = _cast_bias_weight(self, input)
                return self._conv_forward(input, w, b)

        class Embedding(_CastOp, torch.nn.Embedding):
            comfy_cast_weights = cast

            def reset_parameters(self):
                self.bias = None     # comfy.ops gives Embedding a (None) bias so the cast helpers can treat all ops alike
                return None

            def forward_comfy_cast_weights(self, input, out_dtype=None):
                want = out_dtype
                if self.weight.dtype in (torch.float16, torch.bfloat16):
                    out_dtype = None
                w, _ = _cast_bias_weight(self, device=input.device, dtype=out_dtype)
                return torch.nn.functional.embedding(input, w, self.padding_idx, self.max_norm, self.norm_type,
                                                     self.scale_grad_by_freq, self.sparse).to(dtype=want)

        class LayerNorm(_CastOp, torch.nn.LayerNorm):
            comfy_cast_weights = cast

            def forward_comfy_cast_weights(self, input):
                w, b = (None, None) if self.weight is None else _cast_bias_weight(self, input)
                return torch.nn.functional.layer_norm(input, self.normalized_shape, w, b, self.eps)

        class GroupNorm(_CastOp, torch.nn.GroupNorm):
            comfy_cast_weights = cast

            def forward_comfy_cast_weights(self, input):
                w, b = _cast_bias_weight(self, input)
                return torch.nn.functional.group_norm(input, self.num_groups, w, b, self.eps)

        for c in (Linear, Conv2d, Embedding, LayerNorm, GroupNorm):
            ns[c.__name__] = c
        return ns

    disable_weight_init = type("disable_weight_init", (), _family(False))
    manual_cast = type("manual_cast", (disable_weight_init,), _family(True))

    def _calculate_weight(patches, weight, key, intermediate_dtype=torch.float32, original_weights=None):
        for p in patches:
            strength, v, strength_model = p[0], p[1], p[2]
            if strength_model != 1.0:
                weight *= strength_model
            if isinstance(v, torch.Tensor):
                v = ("diff", (v,))
            kind, payload = v[0], v[1]
            if kind == "diff":
                weight += (strength * payload[0].to(weight.device, intermediate_dtype)).to(weight.dtype)
            elif kind == "lora":
                up = payload[0].to(weight.device, intermediate_dtype)
                down = payload[1].to(weight.device, intermediate_dtype)
                scale = 1.0 if payload[2] is None else float(payload[2]) / down.shape[0]
                weight += ((strength * scale) * torch.mm(up.flatten(1), down.flatten(1)).reshape(weight.shape)).to(weight.dtype)
            else:
                raise NotImplementedError(f"stand-in calculate_weight: patch type {kind!r} needs real ComfyUI")
        return weight

    comfy_ops = types.SimpleNamespace(cast_to=_cast_to, cast_bias_weight=_cast_bias_weight, CastWeightBiasOp=_CastOp,
                                      disable_weight_init=disable_weight_init, manual_cast=manual_cast)
    comfy_lora = types.SimpleNamespace(calculate_weight=_calculate_weight)
    comfy_mm = types.SimpleNamespace(
        device_supports_non_blocking=lambda device: torch.device(device).type == "cuda",
        text_encoder_offload_device=lambda: torch.device("cpu"),
    )
