"""Same public names as the reference's `model` package (reference model/__init__.py:1-3)."""
from .criterion import GDL, temporal_weight_func, MSELoss, BiPatchNCE, L1Loss, GANLoss
from .VPTR_modules import VPTREnc, VPTRDec, VPTRDisc, VPTRFormerNAR, VPTRFormerFAR
from .ResNetAutoEncoder import init_weights, clear_packed_weights
