#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" 2>&1 | tail -2
for m in 3; do echo "== mode $m"; SOLR_MODE=$m timeout 300 python tools/gpu/gpu_share_sweep.py config2 config4 2>&1 | grep share; done | tee -a $O/r2A_fused.log
