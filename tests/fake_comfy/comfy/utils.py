"""TEST INFRASTRUCTURE: attribute helpers of comfy.utils used by the patcher."""
import torch


def get_attr(obj, attr):
    for name in attr.split("."):
        obj = getattr(obj, name)
    return obj


def set_attr_param(obj, attr, value):
    *path, leaf = attr.split(".")
    for name in path:
        obj = getattr(obj, name)
    prev = getattr(obj, leaf)
    setattr(obj, leaf, torch.nn.Parameter(value, requires_grad=False))
    return prev


def copy_to_param(obj, attr, value):
    get_attr(obj, attr).data.copy_(value)


def load_torch_file(path, safe_load=True):
    return torch.load(path, weights_only=safe_load)
