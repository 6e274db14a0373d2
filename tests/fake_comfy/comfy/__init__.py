"""Minimal stand-in for the ComfyUI host package (ComfyUI is not installed offline).

Only the symbols the GGUF custom-ops path touches are provided, with signatures
reconstructed from the reference's call sites (ops.py:186-210, 227-271).  Used by
tests/ to run (a) the unmodified reference ops.py when generating golden vectors
and (b) this repo's drop-in ops against the same host interface.
"""
