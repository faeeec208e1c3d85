#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2I_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2I_pytest.log
tail -3 $O/r2I_pytest.log
timeout 300 python tools/gpu/gpu_share_sweep.py config2 2>&1 | grep share | head -2
