"""GPU: the library's debug / A-B environment switches select real code paths (the 1-CTA GEMM, the per-lane GEMM epilogue, the
per-tap implicit conv, the scalar and per-head attention kernels, the non-streaming depthwise conv).  They are read once per
process, so each one is exercised in a subprocess that re-runs the kernel-level parity tests of the family it affects -- a path
kept alive behind a switch is a path that has to stay correct."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GEMM = ["tests/test_gpu_gemm.py"]
ATTN = ["tests/test_gpu_kernels.py", "tests/test_gpu_dropout.py", "-k", "attention"]
MODEL = ["tests/test_gpu_fullsize.py", "-k", "cfg1 or cfg3 or far"]
ENCODER = ["tests/test_gpu_models.py", "tests/test_gpu_fullsize.py", "-k", "autoencoder or packed"]
DWCONV = ["tests/test_gpu_kernels.py", "-k", "dwconv"]

SWITCHES = [
    ("VPTR_GEMM_1CTA", "1", GEMM),            # gemm_tf32_kernel (one CTA per tile) for every shape
    ("VPTR_GEMM_NARROW", "1", GEMM),          # 256 x 176 pair tiles also for the wide (N >= 1024) outputs
    ("VPTR_GEMM_EPI_STG", "1", GEMM),         # per-lane store epilogue instead of TMA bulk stores / bulk-loaded residual
    ("VPTR_CONV_GENERIC", "1", GEMM),         # per-(tap, slice) 4-D TMA box conv instead of the raw-tile kernel (8x8 and quadrant grids)
    ("VPTR_CONV_TF32", "1", ENCODER),         # ResnetBlock convs on the two-plane TF32 raw-tile kernel instead of the bf16x3 one
    ("VPTR_ATTN_TC", "0", MODEL),             # engine keeps the attention forward on the mma.sync kernels (no tcgen05 forward)
    # (VPTR_ATTN_TC=1 makes vptr_attn_fwd itself route to the single-pass TF32 tcgen05 kernel, which changes the numerics to the
    #  tcgen05 tolerance by design; that kernel's generic mask path is covered directly by test_attention_tcgen05_forward)
    ("VPTR_ATTN_NO_MMA", "1", ATTN),          # scalar attention kernels at head_dim 66
    ("VPTR_ATTN_NO_MMA64", "1", ATTN),        # 64-token groups on the scalar kernels
    ("VPTR_ATTN_HPC", "2", ATTN),             # two heads per CTA in attn_mma_kernel
    ("VPTR_ATTN_PERHEAD", "1", ATTN),         # scalar per-head kernels instead of the all-heads ones
    ("VPTR_DWCONV_NOSTREAM", "1", DWCONV),    # register-window depthwise conv instead of the cp.async streaming kernel
]


@pytest.mark.parametrize("name,value,sel", SWITCHES, ids=["%s=%s" % (s[0], s[1]) for s in SWITCHES])
def test_switch_selected_path_passes_its_parity_tests(name, value, sel):
    env = dict(os.environ)
    env[name] = value
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + sel, cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, "%s=%s:\n%s\n%s" % (name, value, tail, r.stderr[-2000:])
    assert " passed" in tail and "failed" not in tail, tail
