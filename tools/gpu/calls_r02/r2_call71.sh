#!/bin/bash
# lanes that hold a primitive wait for company (UW_LEAF_WAIT lanes, at most UW_LEAF_WAIT_MAX iterations): variants w<lanes>_<max> against the default build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2e_leaf_wait.log
timeout 400 python tools/gpu/gpu_lib_sweep.py share > $O 2>&1
grep "^libvar" $O
