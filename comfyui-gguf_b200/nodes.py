"""ComfyUI node classes (drop-in for the reference's nodes.py:134-321).  Importable only inside ComfyUI.

The loader nodes keep their names, widgets and return types; what they hand to ComfyUI is this package's
GGMLOps, so every quantised Linear built by comfy.sd runs the B200 kernels.  GGUFModelPatcher keeps the one
behaviour the hot path depends on: LoRA patches on quantised weights are ATTACHED to the tensor
(`tensor.patches`, nodes.py:43-47) and applied after dequant on every forward instead of being baked in.
"""
import collections
import logging

import torch

import comfy.float
import comfy.lora
import comfy.model_management
import comfy.model_patcher
import comfy.sd
import comfy.utils
import folder_paths
import nodes

from .dequant import is_quantized
from .loader import gguf_clip_loader, gguf_sd_loader
from .ops import GGMLOps, move_patch_to_device

_DTYPE_CHOICES = ["default", "target", "float32", "float16", "bfloat16"]


def update_folder_names_and_paths(key, targets=()):
    """Register a `.gguf`-only file list aliasing an existing model folder (nodes.py:19-32)."""
    known = folder_paths.folder_names_and_paths
    existing = known.get(key, ([], {}))[0]
    existing = existing if isinstance(existing, (list, set, tuple)) else []
    target = next((t for t in targets if t in known), targets[0])
    base_dirs = known.get(target, ([], {}))[0]
    known[key] = (base_dirs or existing, {".gguf"})
    if existing and existing != base_dirs:
        logging.warning(f"Unknown file list already present on key {key}: {existing}")


update_folder_names_and_paths("unet_gguf", ["diffusion_models", "unet"])
update_folder_names_and_paths("clip_gguf", ["text_encoders", "clip"])


class GGUFModelPatcher(comfy.model_patcher.ModelPatcher):
    patch_on_device = False

    def patch_weight_to_device(self, key, device_to=None, inplace_update=False):
        if key not in self.patches:
            return
        weight = comfy.utils.get_attr(self.model, key)
        patches = self.patches[key]
        if is_quantized(weight):
            patched = weight.to(device_to)
            where = self.load_device if self.patch_on_device else self.offload_device
            patched.patches = [(move_patch_to_device(patches, where), key)]
        else:
            inplace_update = self.weight_inplace_update or inplace_update
            if key not in self.backup:
                Backup = collections.namedtuple("Dimension", ["weight", "inplace_update"])
                self.backup[key] = Backup(weight.to(device=self.offload_device, copy=inplace_update), inplace_update)
            if device_to is not None:
                work = comfy.model_management.cast_to_device(weight, device_to, torch.float32, copy=True)
            else:
                work = weight.to(torch.float32, copy=True)
            patched = comfy.lora.calculate_weight(patches, work, key)
            patched = comfy.float.stochastic_rounding(patched, weight.dtype)
        if inplace_update:
            comfy.utils.copy_to_param(self.model, key, patched)
        else:
            comfy.utils.set_attr_param(self.model, key, patched)

    def unpatch_model(self, device_to=None, unpatch_weights=True):
        if unpatch_weights:
            for p in self.model.parameters():
                if is_torch_compatible_param(p):
                    continue
                patches = getattr(p, "patches", [])
                if len(patches) > 0:
                    p.patches = []
        return super().unpatch_model(device_to=device_to, unpatch_weights=unpatch_weights)

    def clone(self, *args, **kwargs):
        src = self
        new = GGUFModelPatcher(src.model, src.load_device, src.offload_device, src.size, weight_inplace_update=src.weight_inplace_update)
        new.patches = {k: v[:] for k, v in src.patches.items()}
        new.patches_uuid = src.patches_uuid
        new.object_patches = src.object_patches.copy()
        new.model_options = __import__("copy").deepcopy(src.model_options)
        new.backup = src.backup
        new.object_patches_backup = src.object_patches_backup
        new.patch_on_device = getattr(src, "patch_on_device", False)
        return new


def is_torch_compatible_param(p):
    return not is_quantized(p)


def _pick_dtype(choice):
    if choice in ("default", None):
        return None
    if choice == "target":
        return "target"
    return getattr(torch, choice)


class UnetLoaderGGUF:
    @classmethod
    def INPUT_TYPES(s):
        return {"required": {"unet_name": (list(folder_paths.get_filename_list("unet_gguf")),)}}

    RETURN_TYPES = ("MODEL",)
    FUNCTION = "load_unet"
    CATEGORY = "bootleg"
    TITLE = "Unet Loader (GGUF)"

    def load_unet(self, unet_name, dequant_dtype=None, patch_dtype=None, patch_on_device=None):
        ops = GGMLOps()
        ops.Linear.dequant_dtype = _pick_dtype(dequant_dtype)   # class attributes on purpose (nodes.py:152-164)
        ops.Linear.patch_dtype = _pick_dtype(patch_dtype)
        unet_path = folder_paths.get_full_path("unet", unet_name)
        sd = gguf_sd_loader(unet_path)
        model = comfy.sd.load_diffusion_model_state_dict(sd, model_options={"custom_operations": ops})
        if model is None:
            logging.error("ERROR UNSUPPORTED UNET {}".format(unet_path))
            raise RuntimeError("ERROR: Could not detect model type of: {}".format(unet_path))
        model = GGUFModelPatcher.clone(model)
        model.patch_on_device = patch_on_device
        return (model,)


class UnetLoaderGGUFAdvanced(UnetLoaderGGUF):
    @classmethod
    def INPUT_TYPES(s):
        return {"required": {
            "unet_name": (list(folder_paths.get_filename_list("unet_gguf")),),
            "dequant_dtype": (_DTYPE_CHOICES, {"default": "default"}),
            "patch_dtype": (_DTYPE_CHOICES, {"default": "default"}),
            "patch_on_device": ("BOOLEAN", {"default": False}),
        }}
    TITLE = "Unet Loader (GGUF/Advanced)"


class CLIPLoaderGGUF:
    N_FILES = 1

    @classmethod
    def INPUT_TYPES(s):
        base = nodes.CLIPLoader.INPUT_TYPES()
        return {"required": {"clip_name": (s.get_filename_list(),), "type": base["required"]["type"]}}

    RETURN_TYPES = ("CLIP",)
    FUNCTION = "load_clip"
    CATEGORY = "bootleg"
    TITLE = "CLIPLoader (GGUF)"

    @classmethod
    def get_filename_list(s):
        return sorted(folder_paths.get_filename_list("clip") + folder_paths.get_filename_list("clip_gguf"))

    def load_data(self, ckpt_paths):
        loaded = []
        for p in ckpt_paths:
            if p.endswith(".gguf"):
                loaded.append(gguf_clip_loader(p))
                continue
            sd = comfy.utils.load_torch_file(p, safe_load=True)
            if "scaled_fp8" in sd:
                raise NotImplementedError(f"Mixing scaled FP8 with GGUF is not supported! Use regular CLIP loader or switch model(s)\n({p})")
            loaded.append(sd)
        return loaded

    def load_patcher(self, clip_paths, clip_type, clip_data):
        clip = comfy.sd.load_text_encoder_state_dicts(
            clip_type=clip_type, state_dicts=clip_data,
            model_options={"custom_operations": GGMLOps, "initial_device": comfy.model_management.text_encoder_offload_device()},
            embedding_directory=folder_paths.get_folder_paths("embeddings"))
        clip.patcher = GGUFModelPatcher.clone(clip.patcher)
        return clip

    def _load(self, names, type):
        paths = tuple(folder_paths.get_full_path("clip", n) for n in names)
        clip_type = getattr(comfy.sd.CLIPType, type.upper(), comfy.sd.CLIPType.STABLE_DIFFUSION)
        return (self.load_patcher(paths, clip_type, self.load_data(paths)),)

    def load_clip(self, clip_name, type="stable_diffusion"):
        return self._load((clip_name,), type)


def _multi_clip(n, title, base_node):
    class Multi(CLIPLoaderGGUF):
        N_FILES = n
        TITLE = title

        @classmethod
        def INPUT_TYPES(s):
            files = (s.get_filename_list(),)
            req = {f"clip_name{i + 1}": files for i in range(n)}
            base = getattr(nodes, base_node, None)
            if base is not None and "type" in base.INPUT_TYPES().get("required", {}):
                req["type"] = base.INPUT_TYPES()["required"]["type"]
            return {"required": req}

        def load_clip(self, *names, type="stable_diffusion", **kw):
            names = list(names) + [kw[f"clip_name{i + 1}"] for i in range(len(names), n)]
            return self._load(tuple(names), type if base_node == "DualCLIPLoader" else "stable_diffusion")
    Multi.__name__ = f"{['', '', 'Dual', 'Triple', 'Quadruple'][n]}CLIPLoaderGGUF"
    return Multi


DualCLIPLoaderGGUF = _multi_clip(2, "DualCLIPLoader (GGUF)", "DualCLIPLoader")
TripleCLIPLoaderGGUF = _multi_clip(3, "TripleCLIPLoader (GGUF)", "TripleCLIPLoader")
QuadrupleCLIPLoaderGGUF = _multi_clip(4, "QuadrupleCLIPLoader (GGUF)", "QuadrupleCLIPLoader")

NODE_CLASS_MAPPINGS = {
    "UnetLoaderGGUF": UnetLoaderGGUF,
    "CLIPLoaderGGUF": CLIPLoaderGGUF,
    "DualCLIPLoaderGGUF": DualCLIPLoaderGGUF,
    "TripleCLIPLoaderGGUF": TripleCLIPLoaderGGUF,
    "QuadrupleCLIPLoaderGGUF": QuadrupleCLIPLoaderGGUF,
    "UnetLoaderGGUFAdvanced": UnetLoaderGGUFAdvanced,
}
