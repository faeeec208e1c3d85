#!/bin/bash
# geometry + primitive records derived on the device at upload: whole GPU suite, upload stage times
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2O_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2O_pytest.log
tail -3 $O/r2O_pytest.log
SOLR_B200_TIMING=1 timeout 600 python tools/gpu/gpu_treebuild.py 2>&1 | tail -22
