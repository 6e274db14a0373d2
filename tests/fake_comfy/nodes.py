"""TEST INFRASTRUCTURE: base classes of ComfyUI's stock CLIP loader nodes."""


class CLIPLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"clip_name": ([],), "type": (["stable_diffusion", "sd3", "flux"],)}}


class DualCLIPLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"clip_name1": ([],), "clip_name2": ([],), "type": (["sdxl", "sd3", "flux"],)}}


class TripleCLIPLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"clip_name1": ([],), "clip_name2": ([],), "clip_name3": ([],)}}


class QuadrupleCLIPLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"clip_name1": ([],), "clip_name2": ([],), "clip_name3": ([],), "clip_name4": ([],)}}
