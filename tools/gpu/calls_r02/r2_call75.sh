#!/bin/bash
# fused stages (option 6 = 2) with release stores / reductions instead of __threadfence(): identical frames, then staged vs fused timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" 2>&1 | tail -2
echo staged; timeout 60 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep "^libvar"
echo fused; SOLR_MODE=2 timeout 60 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep "^libvar"
