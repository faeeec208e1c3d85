#!/bin/bash
# option 12: ownership in blocks of B x B tiles; every rank's 1/8 share in turn on one GPU (the slowest is what an 8-GPU frame takes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for b in 1 2 4 8; do echo "== owner block $b"; SOLR_MODE=1 SOLR_ALL_RANKS=1 SOLR_OWNER_BLOCK=$b timeout 600 python tools/gpu/gpu_share_sweep.py config2 config4 2>&1 | grep "share 1/8" | awk '{print $1, $3, $11, $12}' | tr '\n' ';'; echo; done | tee $O/r2M_owner.log
