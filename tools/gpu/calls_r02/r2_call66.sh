#!/bin/bash
# pushed children prefetched into L2 (p1) / L1 (p2) against the default build, twice
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c_prefetch.log
timeout 300 python tools/gpu/gpu_lib_sweep.py share > $O 2>&1; timeout 300 python tools/gpu/gpu_lib_sweep.py share >> $O 2>&1
grep "^libvar" $O
