#!/bin/bash
# long cylinders split into pieces in the host-built walk trees: timing (config 2 whole frame / share), parity suite on the default
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep part | tee $O/r2F_split_sweep.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/r2F_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2F_pytest.log
tail -3 $O/r2F_pytest.log
