#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -q -x > $O/r2J_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2J_pytest.log
tail -3 $O/r2J_pytest.log
