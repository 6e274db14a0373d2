"""Subset of comfy.lora: `calculate_weight` for "diff" and "lora" patches."""
import torch


def calculate_weight(patches, weight, key, intermediate_dtype=torch.float32, original_weights=None):
    for p in patches:
        strength, v, strength_model = p[0], p[1], p[2]
        if strength_model != 1.0:
            weight *= strength_model
        if isinstance(v, torch.Tensor):
            v = ("diff", (v,))
        kind, payload = v[0], v[1]
        if kind == "diff":
            diff = payload[0]
            weight += (strength * diff.to(device=weight.device, dtype=intermediate_dtype)).to(weight.dtype)
        elif kind == "lora":
            up = payload[0].to(device=weight.device, dtype=intermediate_dtype)
            down = payload[1].to(device=weight.device, dtype=intermediate_dtype)
            alpha = payload[2]
            scale = 1.0 if alpha is None else float(alpha) / down.shape[0]
            delta = torch.mm(up.flatten(start_dim=1), down.flatten(start_dim=1)).reshape(weight.shape)
            weight += ((strength * scale) * delta).to(weight.dtype)
        else:
            raise NotImplementedError(f"fake comfy.lora: patch type {kind!r}")
    return weight
