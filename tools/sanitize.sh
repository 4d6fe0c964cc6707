#!/bin/bash
# compute-sanitizer passes over the GPU tests (run on the GPU box): memcheck on everything, racecheck on the kernels that
# share tiles between warps.  Last run (round 1): memcheck 0 errors (89 tests); racecheck 0 hazards on the attention, depthwise
# conv, LayerNorm, norm+GELU and head kernels (the only reports were inside torch's own layer_norm backward, used by the
# reference side of the tests).
cd "$(dirname "$0")/.."
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "Host Frame\|Device Frame" | tail -5
compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu \
    -k "attention_core or dwconv or layernorm or norm_act or head" 2>&1 | grep -v "Host Frame\|Device Frame" | tail -8
