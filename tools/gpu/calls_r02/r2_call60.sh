#!/bin/bash
# evict-first / no-L1 accesses to write-once-read-once data (STREAM_HINTS 1: parked paths, colours; 2: + the frame's buffers) against the default build, twice
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2W_hints.log
timeout 300 python tools/gpu/gpu_lib_sweep.py share > $O 2>&1
timeout 300 python tools/gpu/gpu_lib_sweep.py share >> $O 2>&1
grep "^libvar" $O
