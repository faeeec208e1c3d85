#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2e_lib_sweep.log 2>&1
cat $O/r2e_lib_sweep.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/r2e_pytest.log 2>&1
echo "pytest rc $?" >> $O/r2e_pytest.log; tail -5 $O/r2e_pytest.log
