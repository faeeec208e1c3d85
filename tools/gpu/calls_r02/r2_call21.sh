#!/bin/bash
# wavefront driver: node visits / primitive tests against the staged driver (debug counters); source-level captures of the walk kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
echo "== mode 1"; SOLR_MODE=1 timeout 300 python tools/gpu/gpu_lib_sweep.py 2>&1 | tail -3
echo "== mode 2"; SOLR_MODE=2 timeout 300 python tools/gpu/gpu_lib_sweep.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 43 --launch-count 5 -o $O/r2u_wave -f python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2u_ncu.log 2>&1
tail -2 $O/r2u_ncu.log
