#!/bin/bash
# compute-sanitizer synccheck over the streamed-output kernels (ballots, match_any and shuffles of streamBatch at the warp's convergent point)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 compute-sanitizer --tool synccheck python -m pytest tests/test_streamed_output_gpu.py -m gpu -q -x > $O/r2d_synccheck.log 2>&1; echo "rc $?" >> $O/r2d_synccheck.log
grep -v "^$" $O/r2d_synccheck.log | tail -6
