#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2c_lib_sweep.log 2>&1
cat $O/r2c_lib_sweep.log
