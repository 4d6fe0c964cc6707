"""ctypes binding of libvptr_b200.so (include/vptr_b200.h).  No CPU fallback: if the library is missing or a call
fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvptr_b200.so")

_lib = None

P = ctypes.c_void_p
I = ctypes.c_int
L = ctypes.c_longlong
F = ctypes.c_float
U = ctypes.c_ulonglong

_SIGS = {
    "vptr_version": ([], I),
    "vptr_gemm_tf32": ([P, L, I, P, L, I, P, L, I, I, I, P, P, L, F, I, I, I, P, I, U, F, P], I),
    "vptr_gemm_debug_buffer": ([P], I),
    "vptr_gemm_simt": ([P, L, I, P, L, I, P, L, I, I, I, P, P, L, F, I, I, I, P, I, U, F, P], I),
    "vptr_layernorm_fwd": ([P, P, P, P, P, P, I, I, P, P, L, I, F, I, I, P], I),
    "vptr_layernorm_bwd": ([P, P, P, P, P, P, P, P, P, P, P, L, I, I, P], I),
    "vptr_bn_stats": ([P, L, I, P, P, P, P, F, F, P, P], I),
    "vptr_bn_eval_stats": ([P, P, P, P, I, F, P], I),
    "vptr_group_stats": ([P, I, L, P, P, F, P], I),
    "vptr_norm_act_fwd": ([P, P, P, P, P, P, P, L, I, I, I, I, P, I, U, F, P], I),
    "vptr_norm_act_bwd": ([P, P, P, P, P, P, P, P, P, L, I, I, I, P, I, P, I, U, F, P], I),
    "vptr_attn_fwd": ([P, L, P, L, P, L, P, L, P, I, I, I, I, I, I, I, I, I, I, F, I, U, F, P], I),
    "vptr_attn_fwd_tcgen05": ([P, L, P, L, P, L, P, L, P, I, I, I, I, I, I, I, I, I, I, F, I, U, F, P], I),
    "vptr_attn_bwd": ([P, L, P, L, P, L, P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, I, I, I, I, I, F, I, U, F, P], I),
    "vptr_attn_bwd_bias": ([P, L, P, L, P, L, P, L, P, L, P, L, P, L, P, P, I, I, I, I, I, I, I, I, I, I, F, I, U, F, P, P, P, P], I),
    "vptr_window_index_maps": ([I, I, I, I, P, P, P], I),
    "vptr_causal_mask": ([I, P, P], I),
    "vptr_dwconv3x3": ([P, P, P, P, I, I, I, I, I, P], I),
    "vptr_dwconv3x3_stats": ([P, P, P, P, I, I, I, I, P, P], I),
    "vptr_group_stats_finalize": ([P, I, L, P, P, F, P], I),
    "vptr_dwconv3x3_wgrad": ([P, P, P, P, I, I, I, I, P], I),
    "vptr_axpby": ([P, P, P, L, F, F, P], I),
    "vptr_add_rows": ([P, P, P, L, I, I, I, I, P], I),
    "vptr_rowgroup_sum": ([P, P, L, I, P], I),
    "vptr_gelu_fwd": ([P, P, L, I, U, F, P], I),
    "vptr_gelu_bwd": ([P, P, P, L, I, U, F, P], I),
    "vptr_round_copy": ([P, P, L, I, P, L, U, F, P], I),
    "vptr_round_copy_multi": ([P, I, P, L, P], I),
    "vptr_droppath_scales": ([P, I, U, F, P], I),
    "vptr_relu_fwd": ([P, P, L, P], I),
    "vptr_relu_bwd": ([P, P, P, L, P], I),
    "vptr_colsum": ([P, P, L, I, L, P], I),
    "vptr_transpose": ([P, P, I, I, I, I, P], I),
    "vptr_transpose_multi": ([P, I, I, I, P], I),
    "vptr_pad_crop": ([P, P, I, I, I, I, I, I, I, I, I, P], I),
    "vptr_sqnorm_accumulate": ([P, L, P, P], I),
    "vptr_round_copy_colsum": ([P, P, L, I, I, P, I, U, F, P, P], I),
    "vptr_gelu_bwd_colsum": ([P, P, P, L, I, I, U, F, P, P], I),
    "vptr_norm_act_bwd_colsum": ([P, P, P, P, P, P, P, P, P, L, I, I, I, P, I, P, I, U, F, P, P], I),
    "vptr_clip_scale": ([P, L, P, F, P], I),
    "vptr_conv3x3_tf32": ([P, P, P, I, I, I, I, I, P, P, I, I, I, P], I),
    "vptr_conv3x3_tf32_quad": ([P, P, P, I, I, I, I, I, P, P, I, I, I, P], I),
    "vptr_pad_nhwc_quad": ([P, P, I, I, I, I, I, I, P], I),
    "vptr_conv3x3_bf16x3": ([P, P, P, I, I, I, I, I, P, P, I, I, P], I),
    "vptr_pad_nhwc_quad_bf16x2": ([P, P, I, I, I, I, I, P], I),
    "vptr_split_bf16x2": ([P, P, L, L, P], I),
    "vptr_split_tf32": ([P, P, L, L, P], I),
    "vptr_pad_nhwc": ([P, P, I, I, I, I, I, I, I, P], I),
    "vptr_im2col": ([P, P, P, I, I, I, I, I, I, I, I, I, P], I),
    "vptr_convT_gather": ([P, P, P, I, I, I, I, I, P], I),
    "vptr_bn_fold": ([P, P, P, P, F, P, P, I, P], I),
    "vptr_pack_conv_weight": ([P, P, P, I, I, I, I, P], I),
    "vptr_stem_conv7x7": ([P, P, P, P, I, I, I, I, I, P], I),
    "vptr_head_conv7x7_fwd": ([P, P, P, P, I, I, I, I, I, I, P], I),
    "vptr_head_conv7x7_bwd": ([P, P, P, P, I, I, I, I, I, I, P, P], I),
    "vptr_mse_gdl_fwd": ([P, P, L, I, I, P, P, P], I),
    "vptr_mse_gdl_bwd": ([P, P, P, P, L, I, I, P], I),
    "vptr_sqnorm_multi": ([P, I, L, I, P, P], I),
    "vptr_bipatch_nce_fwd": ([P, P, I, I, I, F, P, P, P, P, P], I),
    "vptr_bipatch_nce_bwd": ([P, P, P, P, P, I, I, I, F, P, P, P], I),
    "vptr_adamw_multi": ([P, I, L, I, F, F, F, F, F, L, P, F, P], I),
    "vptr_stem_conv7x7_raw": ([P, P, P, I, I, I, I, I, P], I),
    "vptr_bn_act_fwd": ([P, P, P, P, P, P, P, L, I, I, I, P], I),
    "vptr_bn_act_bwd": ([P, P, P, P, P, P, P, P, P, P, L, I, I, P, I, P], I),
    "vptr_col2im": ([P, P, I, I, I, I, I, I, I, I, P], I),
    "vptr_stem_wgrad": ([P, P, P, I, I, I, I, P], I),
    "vptr_act_bwd": ([P, P, P, L, I, P], I),
    "vptr_rng_advance": ([L, P], I),
    "vptr_counter_add": ([P, L, P], I),
    "vptr_adamw_multi_dev": ([P, I, L, I, F, F, F, F, F, L, P, P, F, P], I),
    "vptr_nccl_unique_id": ([P], I),
    "vptr_nccl_comm_init": ([P, I, I, P], I),
    "vptr_nccl_comm_destroy": ([P], I),
    "vptr_allreduce_grads": ([P, P, L, P, P], I),
    "vptr_workspace_bytes": ([I, L, I, I, I], L),
}

EXPORTS = tuple(_SIGS) + ("vptr_last_error",)

launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("vptr_b200: %s is missing -- build it with `python -m vptr_b200.build` "
                               "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = res
        l.vptr_last_error.argtypes = []
        l.vptr_last_error.restype = ctypes.c_char_p
        _lib = l
    return _lib


def call(name, *args):
    """Invoke an entry point; non-zero status -> RuntimeError carrying vptr_last_error()."""
    global launch_count
    l = lib()
    rc = getattr(l, name)(*args)
    launch_count += 1
    if rc != 0:
        raise RuntimeError("%s failed (status %d): %s" % (name, rc, l.vptr_last_error().decode("utf-8", "replace")))
