#!/bin/bash
# fused stages (option 6 = 3): frames identical?  timing against the staged driver, whole frame and 1/8 share, config 2 and config 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" > $O/r2A_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2A_pytest.log
tail -5 $O/r2A_pytest.log
for m in 1 3; do echo "== mode $m"; SOLR_MODE=$m timeout 300 python tools/gpu/gpu_share_sweep.py config2 config4 2>&1 | grep share; done | tee $O/r2A_fused.log
