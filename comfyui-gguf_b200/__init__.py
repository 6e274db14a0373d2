"""ggufb200: B200-native (sm_100a) drop-in for the dequant + Linear hot path of city96/ComfyUI-GGUF.

Loaded by ComfyUI as a custom node directory (node classes exported when `comfy` is importable), or by path
through `__graft_entry__.load_package()` for tests and benchmarks.
"""
try:
    import comfy.utils  # noqa: F401
    import folder_paths  # noqa: F401
except ImportError:
    pass
else:  # pragma: no cover - only inside ComfyUI
    from .nodes import NODE_CLASS_MAPPINGS
    NODE_DISPLAY_NAME_MAPPINGS = {k: v.TITLE for k, v in NODE_CLASS_MAPPINGS.items()}
    __all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
