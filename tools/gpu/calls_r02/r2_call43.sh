#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 1 2; do timeout 600 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep part; done
