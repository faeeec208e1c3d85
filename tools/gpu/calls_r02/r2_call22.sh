#!/bin/bash
# gather replay fast path in the unit walk: parity suite, timing (whole frame and 1/8 share)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2v_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2v_pytest.log
tail -5 $O/r2v_pytest.log
timeout 300 python tools/gpu/gpu_lib_sweep.py share 2>&1 | tail -3
