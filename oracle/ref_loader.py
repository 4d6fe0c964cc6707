"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference under /root/reference.

Only `tests/golden/make_golden.py` (run in the build container, where /root/reference
exists) and `bench.py --impl reference` (when the tree is present) may use this.  Nothing
in the product package `vptr_b200/` imports it.

The reference needs two non-invasive adapters to import on this image:
  * `timm` is absent: a 2-symbol `timm.models.layers` shim (`to_2tuple`, `trunc_normal_`;
    used at model/VidHRFormer_modules.py:4 and model/MultiHeadAttentionRPE.py:19).
  * utils/position_encoding.py:56,100 default to device cuda:0; VPTR_modules.py:123,127,179
    call them with defaults -> rebind to functools.partial(..., device=dev).
"""
import functools
import importlib
import sys
import types

import torch

REF_ROOT = "/root/reference"


def _install_timm_shim():
    if "timm.models.layers" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    def to_2tuple(x):
        if isinstance(x, (tuple, list)):
            return tuple(x)
        return (x, x)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = trunc_normal_
    timm.models = models
    models.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.models"] = models
    sys.modules["timm.models.layers"] = layers


def _install_optional_stubs():
    """utils/__init__.py pulls in dataset.py (cv2, torchvision, PIL) and train_summary.py
    (tensorboard, PIL).  Stub whichever is missing; none of it is on the hot path."""
    for name in ("cv2",):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)


def load_reference(device="cpu", root=REF_ROOT):
    """Returns the reference's `model` package with pos-embedding devices rebound."""
    _install_timm_shim()
    _install_optional_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    for k in list(sys.modules):
        if k == "model" or k.startswith("model.") or k == "utils" or k.startswith("utils."):
            mod = sys.modules[k]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(root):
                del sys.modules[k]
    model = importlib.import_module("model")
    vm = importlib.import_module("model.VPTR_modules")
    pe = importlib.import_module("utils.position_encoding")
    dev = torch.device(device)
    vm.PositionEmbeddding2D = functools.partial(pe.PositionEmbeddding2D, device=dev)
    vm.PositionEmbeddding3D = functools.partial(pe.PositionEmbeddding3D, device=dev)
    return model
