#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python tools/gpu/gpu_splitdiff.py 2>&1 | grep " vs "
