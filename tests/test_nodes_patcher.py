"""CPU tests of nodes.py::GGUFModelPatcher against the ModelPatcher behaviours the reference relies on (nodes.py:43-132),
using tests/fake_comfy (TEST INFRASTRUCTURE stand-ins for comfy.model_patcher / comfy.utils / folder_paths / nodes)."""
import importlib

import numpy as np
import pytest
import torch
import gguf

import oracle
from util import Q


@pytest.fixture(scope="module")
def nodes_mod(pkg):
    import __graft_entry__ as ge
    ge.load_package()
    return importlib.import_module(f"{ge.PKG_NAME}.nodes")


def _model(pkg):
    """Two quantised Linears and one dense one behind the drop-in ops."""
    ops = pkg.ops.GGMLOps

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = ops.Linear(256, 32)
            self.b = ops.Linear(256, 32)
            self.c = torch.nn.Linear(8, 8, bias=False)

    net = Net()
    for i, name in enumerate(("a", "b")):
        raw = oracle.random_blocks(int(Q.Q4_K), 32, seed=i).reshape(32, 144)
        w = pkg.ops.GGMLTensor(torch.from_numpy(raw), tensor_type=Q.Q4_K, tensor_shape=torch.Size((32, 256)))
        getattr(net, name).load_state_dict({"weight": w})
    return net


def _lora(n, k, r=4):
    g = torch.Generator().manual_seed(n + k)
    return ("lora", (torch.randn(n, r, generator=g), torch.randn(r, k, generator=g), 2.0, None, None, None))


def test_node_table_matches_reference_keys(nodes_mod):
    assert set(nodes_mod.NODE_CLASS_MAPPINGS) == {"UnetLoaderGGUF", "CLIPLoaderGGUF", "DualCLIPLoaderGGUF", "TripleCLIPLoaderGGUF",
                                                  "QuadrupleCLIPLoaderGGUF", "UnetLoaderGGUFAdvanced"}      # nodes.py:314-321
    for cls in nodes_mod.NODE_CLASS_MAPPINGS.values():
        assert cls.CATEGORY == "bootleg" and hasattr(cls, "TITLE") and hasattr(cls, "FUNCTION")


def test_lowvram_module_still_gets_tensor_patches(pkg, nodes_mod):
    """without the load() override a partially loaded module would carry LowVramPatch objects that the GGUF layers
    never read -- the LoRA would be silently ignored.  With it every patched key goes through patch_weight_to_device."""
    net = _model(pkg)
    cpu = torch.device("cpu")
    patcher = nodes_mod.GGUFModelPatcher(net, cpu, cpu)
    patcher.add_patches({"a.weight": _lora(32, 256), "b.weight": _lora(32, 256)})
    one_layer = net.a.weight.numel() * net.a.weight.element_size()
    patcher.load(cpu, lowvram_model_memory=one_layer)              # `b` does not fit: stays offloaded
    for layer in (net.a, net.b):
        assert not getattr(layer, "weight_function", None), "a LowVramPatch would never be applied by GGMLLayer"
        assert len(layer.weight.patches) == 1 and layer.weight.patches[0][1].endswith(".weight")
        assert pkg.dequant.is_quantized(layer.weight)                    # the packed bytes are untouched
    terms = net.b._lora_terms(cpu)
    assert terms and abs(terms[0][0] - 2.0 / 4) < 1e-12                 # recognised as a plain LoRA: scale = alpha / rank
    patcher.unpatch_model()
    assert net.a.weight.patches == [] and net.b.weight.patches == []


def test_clone_goes_through_the_base_class(pkg, nodes_mod):
    import comfy.model_patcher
    net = _model(pkg)
    cpu = torch.device("cpu")
    stock = comfy.model_patcher.ModelPatcher(net, cpu, cpu, size=123)
    stock.callbacks = {"on_load": [len]}
    stock.wrappers = {"outer": [str]}
    stock.hook_mode = "min_vram"
    stock.add_patches({"a.weight": _lora(32, 256)})
    twin = nodes_mod.GGUFModelPatcher.clone(stock)                       # how the loader nodes call it (nodes.py:175)
    assert type(twin) is nodes_mod.GGUFModelPatcher and type(stock) is comfy.model_patcher.ModelPatcher
    assert twin.callbacks == stock.callbacks and twin.wrappers == stock.wrappers and twin.hook_mode == "min_vram"
    assert twin.patches.keys() == stock.patches.keys() and twin.patches is not stock.patches
    assert twin.size == 0                                                # foreign source class: size is recomputed
    twin.patch_on_device = True
    again = twin.clone()
    assert type(again) is nodes_mod.GGUFModelPatcher and again.patch_on_device is True and again.size == twin.size
