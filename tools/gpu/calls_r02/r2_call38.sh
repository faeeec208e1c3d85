#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_integration.py -m gpu -q -x > $O/r2E_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2E_pytest.log
tail -25 $O/r2E_pytest.log
