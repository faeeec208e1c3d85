#!/bin/bash
# where an end-to-end frame's time goes, outputs copied / streamed
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -6
timeout 300 python tools/gpu/gpu_stream_e2e.py config4 2>&1 | tail -6
