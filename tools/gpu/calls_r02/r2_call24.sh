#!/bin/bash
# GPU-built walk trees: frames identical to the host-built ones?  upload time and frame time on config 2 (and the 1 M scenes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gpu_built" > $O/r2x_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2x_pytest.log
tail -15 $O/r2x_pytest.log
timeout 900 python tools/gpu/gpu_treebuild.py big 2>&1 | tee $O/r2x_treebuild.log | tail -12
