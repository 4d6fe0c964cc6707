"""TEST / BASELINE INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (XiYe20/VPTR).

The tree is looked up at /root/reference (build container) or at oracle/_ref/ (a byte-for-byte copy staged by
oracle/make_ref.sh; git-ignored, travels to the GPU box with the gpurun snapshot).  Users: tests/golden/make_golden.py,
tests/test_gpu_dropin.py, tests/test_reference_arm.py and bench.py's `--impl reference` / `cpu_baseline` legs.  Nothing in
the product package `vptr_b200/` imports this file.

The reference needs two non-invasive adapters to import on this image:
  * `timm` is absent: a 2-symbol `timm.models.layers` shim (`to_2tuple`, `trunc_normal_`;
    used at model/VidHRFormer_modules.py:4 and model/MultiHeadAttentionRPE.py:19).
  * utils/position_encoding.py:56,100 default to device cuda:0; VPTR_modules.py:123,127,179
    call them with defaults -> rebind to functools.partial(..., device=dev) when running the reference on the CPU.
and its training scripts read their hyper-parameters from globals defined under `if __name__ == '__main__'`
(train_NAR.py:34-36,85), which `load_train_script` injects with setattr.
"""
import functools
import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
STAGED = os.path.join(_HERE, "_ref")


def ref_root():
    """directory of the unmodified reference tree, or None"""
    for r in ("/root/reference", STAGED):
        if os.path.isdir(os.path.join(r, "model")) and os.path.isfile(os.path.join(r, "train_NAR.py")):
            return r
    return None


REF_ROOT = ref_root()


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


_OURS = ("model", "utils", "train_NAR", "train_FAR", "train_NAR_mp", "train_FAR_mp", "train_AutoEncoder")


def _purge():
    for k in list(sys.modules):
        if k in _OURS or k.split(".")[0] in ("model", "utils"):
            del sys.modules[k]


def _set_path(root, dropin):
    drop = os.path.join(REPO_ROOT, "vptr_b200")
    sys.path[:] = [p for p in sys.path if p not in (root, drop)]
    sys.path.insert(0, root)
    if dropin:   # `import model` now resolves to vptr_b200/model; `utils`, train_*.py stay the reference's
        sys.path.insert(0, drop)


def load_reference(device="cpu", root=None):
    """Returns the reference's `model` package with pos-embedding devices rebound."""
    root = root or ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (neither /root/reference nor oracle/_ref; run oracle/make_ref.sh)")
    _install_timm_shim()
    _install_optional_stubs()
    _purge()
    _set_path(root, dropin=False)
    model = importlib.import_module("model")
    vm = importlib.import_module("model.VPTR_modules")
    pe = importlib.import_module("utils.position_encoding")
    dev = torch.device(device)
    vm.PositionEmbeddding2D = functools.partial(pe.PositionEmbeddding2D, device=dev)
    vm.PositionEmbeddding3D = functools.partial(pe.PositionEmbeddding3D, device=dev)
    return model


def load_train_script(name, dropin, device="cpu", root=None, **globals_):
    """Imports the reference's unmodified train_NAR / train_FAR module and injects the `__main__` globals its single_iter /
    cal_lossT read.  dropin=False: `model` is the reference's own package (the CPU baseline).  dropin=True: `model` resolves
    to vptr_b200/model -- the drop-in claim under test.  Returns (train_module, model_package)."""
    root = root or ref_root()
    if root is None:
        raise RuntimeError("reference tree not found (neither /root/reference nor oracle/_ref; run oracle/make_ref.sh)")
    if dropin:
        _install_timm_shim()
        _install_optional_stubs()
        _purge()
        _set_path(root, dropin=True)
        model = importlib.import_module("model")
    else:
        model = load_reference(device, root)
    mod = importlib.import_module(name)
    f = os.path.abspath(mod.__file__)
    if not f.startswith(os.path.abspath(root)):
        raise RuntimeError("%s resolved to %s, not the reference tree" % (name, f))
    for k, v in globals_.items():
        setattr(mod, k, v)
    return mod, model


def unload():
    """drops the reference's (or the drop-in's) top-level `model` / `utils` / train_* modules and path entries again"""
    _purge()
    root = ref_root()
    drop = os.path.join(REPO_ROOT, "vptr_b200")
    sys.path[:] = [p for p in sys.path if p not in (root, drop)]
