#!/bin/bash
# option 11: 32 / 16 / 8 pixels or queue entries per warp and work item; whole frame and 1/8 share, config 2 and config 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for b in 0 1 2; do echo "== batch shift $b"; SOLR_MODE=1 SOLR_BATCH=$b timeout 300 python tools/gpu/gpu_share_sweep.py config2 config4 2>&1 | grep share; done | tee $O/r2L_batch.log
