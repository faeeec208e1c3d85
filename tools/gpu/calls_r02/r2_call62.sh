#!/bin/bash
# full capture of ONE frame's four ray kernels (second frame of the run; kernel-name filter, so that upload kernels do not shift the window)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage --launch-skip 4 --launch-count 4 -o $O/r2Y_frame -f python tools/gpu/prof_staged.py 1 4 3 3 > $O/r2Y_ncu.log 2>&1
tail -2 $O/r2Y_ncu.log
