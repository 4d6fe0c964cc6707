"""Same public names as the reference's `model` package (reference model/__init__.py:1-3).

Importable two ways: as `vptr_b200.model`, or -- with `vptr_b200/` itself on sys.path, which is how the reference's unmodified
train_NAR.py / train_FAR.py (`from model import VPTREnc, ...`, train_NAR.py:13-14) pick it up -- as the top-level package
`model`.  In the second case this module aliases itself to `vptr_b200.model` so both names share one set of classes."""
if __name__ != "vptr_b200.model":
    import importlib
    import os
    import sys
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.append(_root)
    _real = importlib.import_module("vptr_b200.model")
    sys.modules[__name__] = _real
    for _sub in ("criterion", "VPTR_modules", "ResNetAutoEncoder"):
        sys.modules[__name__ + "." + _sub] = sys.modules["vptr_b200.model." + _sub]
else:
    from .criterion import GDL, temporal_weight_func, MSELoss, BiPatchNCE, L1Loss, GANLoss
    from .VPTR_modules import VPTREnc, VPTRDec, VPTRDisc, VPTRFormerNAR, VPTRFormerFAR
    from .ResNetAutoEncoder import init_weights, clear_packed_weights
