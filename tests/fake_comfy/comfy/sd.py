"""TEST INFRASTRUCTURE: names the loader nodes import; the node bodies are not exercised in the CPU suite."""
import enum


class CLIPType(enum.Enum):
    STABLE_DIFFUSION = 1
    SD3 = 3
    FLUX = 6


def load_diffusion_model_state_dict(sd, model_options=None):
    raise NotImplementedError


def load_text_encoder_state_dicts(**kwargs):
    raise NotImplementedError
