#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python tools/gpu/gpu_lib_sweep.py 2>&1 | tail -8
