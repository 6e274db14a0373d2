import torch


def device_supports_non_blocking(device):
    return torch.device(device).type == "cuda"


def get_torch_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def text_encoder_offload_device():
    return torch.device("cpu")


def cast_to_device(tensor, device, dtype, copy=False):
    return tensor.to(device=device, dtype=dtype, copy=copy)
