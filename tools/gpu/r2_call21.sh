#!/bin/bash
# wavefront driver: source-level captures of the closest-hit walk kernels of pass 0 and pass 1, and of the final shade kernel of pass 0
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 43 --launch-count 5 -o $O/r2u_wave -f python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2u_ncu.log 2>&1
tail -2 $O/r2u_ncu.log
