"""TEST INFRASTRUCTURE: the slice of comfy.model_patcher.ModelPatcher that the GGUF patcher subclass touches
(patch bookkeeping, clone, load with the lowvram / force_patch_weights switch, unpatch).  Behaviour modelled on ComfyUI's
public semantics: a partially loaded module gets a LowVramPatch in `weight_function` unless force_patch_weights is set."""
import copy
import uuid

import torch

import comfy.utils


class LowVramPatch:
    def __init__(self, key, patches):
        self.key, self.patches = key, patches


class ModelPatcher:
    def __init__(self, model, load_device, offload_device, size=0, weight_inplace_update=False):
        self.model = model
        self.load_device, self.offload_device = load_device, offload_device
        self.size = size
        self.weight_inplace_update = weight_inplace_update
        self.patches = {}
        self.backup = {}
        self.object_patches = {}
        self.object_patches_backup = {}
        self.model_options = {"transformer_options": {}}
        self.patches_uuid = uuid.uuid4()
        # state a hand-written clone() tends to forget
        self.callbacks = {}
        self.wrappers = {}
        self.hook_mode = "default"

    def model_size(self):
        if self.size == 0:
            self.size = sum(p.numel() * p.element_size() for p in self.model.parameters())
        return self.size

    def clone(self):
        n = self.__class__(self.model, self.load_device, self.offload_device, self.size, weight_inplace_update=self.weight_inplace_update)
        n.patches = {k: v[:] for k, v in self.patches.items()}
        n.patches_uuid = self.patches_uuid
        n.object_patches = self.object_patches.copy()
        n.model_options = copy.deepcopy(self.model_options)
        n.backup = self.backup
        n.object_patches_backup = self.object_patches_backup
        n.callbacks = {k: list(v) for k, v in self.callbacks.items()}
        n.wrappers = {k: list(v) for k, v in self.wrappers.items()}
        n.hook_mode = self.hook_mode
        return n

    def add_patches(self, patches, strength_patch=1.0, strength_model=1.0):
        for key, value in patches.items():
            self.patches.setdefault(key, []).append((strength_patch, value, strength_model, None, None))
        self.patches_uuid = uuid.uuid4()
        return list(patches)

    def patch_weight_to_device(self, key, device_to=None, inplace_update=False):
        raise NotImplementedError("the GGUF subclass overrides this")

    def load(self, device_to=None, lowvram_model_memory=0, force_patch_weights=False, full_load=False):
        """Modules are 'loaded' in registration order until lowvram_model_memory bytes are used; the rest stays offloaded."""
        used = 0
        for name, module in self.model.named_modules():
            if not hasattr(module, "weight") or module.weight is None:
                continue
            key = f"{name}.weight"
            nbytes = module.weight.numel() * module.weight.element_size()
            lowvram = lowvram_model_memory > 0 and not full_load and used + nbytes > lowvram_model_memory
            if lowvram:
                if key in self.patches:
                    if force_patch_weights:
                        self.patch_weight_to_device(key)                      # stays on the offload device
                    else:
                        module.weight_function = [LowVramPatch(key, self.patches)]
                continue
            used += nbytes
            if key in self.patches:
                self.patch_weight_to_device(key, device_to)
            else:
                comfy.utils.set_attr_param(self.model, key, comfy.utils.get_attr(self.model, key).to(device_to))

    def unpatch_model(self, device_to=None, unpatch_weights=True):
        for module in self.model.modules():
            if getattr(module, "weight_function", None):
                module.weight_function = []
