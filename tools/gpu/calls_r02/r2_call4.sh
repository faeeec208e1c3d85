#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
python tools/gpu/prof_staged.py 1 4 3 4 > $O/r2f_time.log 2>&1
# one frame's kernels under ncu --set full: skip the first frame (3 ray kernels + reflected = 4 launches per frame)
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 3 -o $O/r2f_lean_frame -f python tools/gpu/prof_staged.py 1 4 3 2 > $O/r2f_ncu.log 2>&1
tail -3 $O/r2f_time.log; tail -3 $O/r2f_ncu.log
