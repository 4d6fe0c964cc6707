"""vptr_b200 -- B200-native (sm_100a) implementation of VPTR's stage-2 training hot path behind the reference's
nn.Module API.  `vptr_b200.model` mirrors the reference's `model` package; `libvptr_b200.so` holds the kernels."""
__all__ = ["model", "ops", "engine"]
