#!/bin/bash
# child boxes as centre + half extent in the unit walk (no sign selects in the node round): A/B against the lo/hi form, then the whole GPU suite (exactness)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python tools/gpu/gpu_lib_sweep.py share > $O/r2Z_centre_half.log 2>&1; timeout 300 python tools/gpu/gpu_lib_sweep.py share >> $O/r2Z_centre_half.log 2>&1
grep "^libvar" $O/r2Z_centre_half.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2Z_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2Z_pytest.log
tail -4 $O/r2Z_pytest.log
