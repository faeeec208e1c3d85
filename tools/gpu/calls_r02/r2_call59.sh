#!/bin/bash
# L2 experiments: evict-first hints on the parked paths (build variant), a persisting-L2 window over the walk-tree nodes / the primitive records
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2V_l2.log
echo "== plain" > $O; timeout 300 python tools/gpu/gpu_lib_sweep.py share >> $O 2>&1
for mb in 48 80; do for what in 0 1; do echo "== persist $mb MB what $what" >> $O; SOLR_B200_L2_PERSIST=$mb SOLR_B200_L2_PERSIST_WHAT=$what timeout 300 python tools/gpu/gpu_lib_sweep.py share >> $O 2>&1; done; done
grep "^==\|^libvar\|persisting L2" $O | uniq
