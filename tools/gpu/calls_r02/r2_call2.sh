#!/bin/bash
# Round 2, GPU call 2: the lean node round — exactness (GPU parity suite on the default build) and timing of build variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -q -x > $O/r2b_pytest.log 2>&1
echo "pytest rc $?" >> $O/r2b_pytest.log
timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2b_lib_sweep.log 2>&1
tail -4 $O/r2b_pytest.log; cat $O/r2b_lib_sweep.log
