#!/bin/bash
# source-level capture of the staged driver's three ray kernels (default build): where do the warp instructions go, at how many lanes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 8 --launch-count 3 -o $O/r2q_unit_frame -f python tools/gpu/prof_staged.py 1 4 3 4 > $O/r2q_ncu.log 2>&1
tail -3 $O/r2q_ncu.log
