#!/bin/bash
# 4-GPU box: which hoisted half breaks the 3-GPU graph placement at 32x32?
T="tests/test_multigpu.py::test_encrypted_forward_three_gpu_placement_equals_single_gpu"
run() { echo "== $*"; env "$@" timeout 150 python -m pytest "$T" -m gpu -q -k "True" 2>&1 | grep -E "passed|failed|Error|error:|illegal" | head -4; }
run PRIMIA_HOIST_WEIGHT_SIDE=0
run PRIMIA_HOIST_PARTS=conv
run PRIMIA_HOIST_PARTS=newton
run PRIMIA_HOIST_PARTS=newton,bn
run PRIMIA_HOIST_PARTS=newton,conv,bn PRIMIA_FUSE_OPEN=0
