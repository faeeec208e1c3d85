#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "device_animation" > $O/r2C_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2C_pytest.log
tail -30 $O/r2C_pytest.log
