#!/bin/bash
# Round 2, GPU call 10: TMA variants of the unit walk (L2 bulk prefetch of the scene arrays, tree top staged in shared memory), group walk with the leaf batch out of line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2j_lib_sweep.log 2>&1
cat $O/r2j_lib_sweep.log
