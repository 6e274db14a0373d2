"""Subset of comfy.ops: cast helpers and the disable_weight_init / manual_cast op families."""
import torch


def cast_to(weight, dtype=None, device=None, non_blocking=False, copy=False):
    if weight is None:
        return None
    if device is None or weight.device == torch.device(device):
        if not copy and (dtype is None or weight.dtype == dtype):
            return weight
        return weight.to(dtype=dtype, copy=copy)
    return weight.to(device=device, dtype=dtype, non_blocking=non_blocking, copy=copy)


def cast_bias_weight(s, input=None, dtype=None, device=None, bias_dtype=None):
    if input is not None:
        dtype = input.dtype if dtype is None else dtype
        bias_dtype = dtype if bias_dtype is None else bias_dtype
        device = input.device if device is None else device
    bias = cast_to(s.bias, bias_dtype, device) if getattr(s, "bias", None) is not None else None
    weight = cast_to(s.weight, dtype, device)
    return weight, bias


class CastWeightBiasOp:
    comfy_cast_weights = False
    weight_function = []
    bias_function = []


def _make_family(cast_default):
    class Family:
        class Linear(torch.nn.Linear, CastWeightBiasOp):
            comfy_cast_weights = cast_default

            def reset_parameters(self):
                return None

            def forward_comfy_cast_weights(self, input):
                weight, bias = cast_bias_weight(self, input)
                return torch.nn.functional.linear(input, weight, bias)

            def forward(self, *args, **kwargs):
                if self.comfy_cast_weights or len(self.weight_function) > 0 or len(self.bias_function) > 0:
                    return self.forward_comfy_cast_weights(*args, **kwargs)
                return super().forward(*args, **kwargs)

        class Conv2d(torch.nn.Conv2d, CastWeightBiasOp):
            comfy_cast_weights = cast_default

            def reset_parameters(self):
                return None

            def forward_comfy_cast_weights(self, input):
                weight, bias = cast_bias_weight(self, input)
                return self._conv_forward(input, weight, bias)

            def forward(self, *args, **kwargs):
                if self.comfy_cast_weights or len(self.weight_function) > 0 or len(self.bias_function) > 0:
                    return self.forward_comfy_cast_weights(*args, **kwargs)
                return super().forward(*args, **kwargs)

        class Embedding(torch.nn.Embedding, CastWeightBiasOp):
            comfy_cast_weights = cast_default

            def reset_parameters(self):
                self.bias = None
                return None

            def forward_comfy_cast_weights(self, input, out_dtype=None):
                output_dtype = out_dtype
                if self.weight.dtype in (torch.float16, torch.bfloat16):
                    out_dtype = None
                weight, _ = cast_bias_weight(self, device=input.device, dtype=out_dtype)
                return torch.nn.functional.embedding(input, weight, self.padding_idx, self.max_norm, self.norm_type,
                                                     self.scale_grad_by_freq, self.sparse).to(dtype=output_dtype)

            def forward(self, *args, **kwargs):
                if self.comfy_cast_weights or len(self.weight_function) > 0 or len(self.bias_function) > 0:
                    return self.forward_comfy_cast_weights(*args, **kwargs)
                kwargs.pop("out_dtype", None)
                return super().forward(*args, **kwargs)

        class LayerNorm(torch.nn.LayerNorm, CastWeightBiasOp):
            comfy_cast_weights = cast_default

            def reset_parameters(self):
                return None

            def forward_comfy_cast_weights(self, input):
                if self.weight is not None:
                    weight, bias = cast_bias_weight(self, input)
                else:
                    weight, bias = None, None
                return torch.nn.functional.layer_norm(input, self.normalized_shape, weight, bias, self.eps)

            def forward(self, *args, **kwargs):
                if self.comfy_cast_weights or len(self.weight_function) > 0 or len(self.bias_function) > 0:
                    return self.forward_comfy_cast_weights(*args, **kwargs)
                return super().forward(*args, **kwargs)

        class GroupNorm(torch.nn.GroupNorm, CastWeightBiasOp):
            comfy_cast_weights = cast_default

            def reset_parameters(self):
                return None

            def forward_comfy_cast_weights(self, input):
                weight, bias = cast_bias_weight(self, input)
                return torch.nn.functional.group_norm(input, self.num_groups, weight, bias, self.eps)

            def forward(self, *args, **kwargs):
                if self.comfy_cast_weights or len(self.weight_function) > 0 or len(self.bias_function) > 0:
                    return self.forward_comfy_cast_weights(*args, **kwargs)
                return super().forward(*args, **kwargs)

    return Family


disable_weight_init = _make_family(False)
disable_weight_init.__name__ = "disable_weight_init"


class manual_cast(_make_family(True)):
    pass
