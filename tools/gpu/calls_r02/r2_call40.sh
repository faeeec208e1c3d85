#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep part
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2G_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2G_pytest.log
tail -3 $O/r2G_pytest.log
