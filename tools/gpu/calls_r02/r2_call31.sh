#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python tools/gpu/gpu_share_sweep.py config4 config2 2>&1 | grep "share" | tee gpurun_out/r2_share_sweep.log
