#!/bin/bash
# streamBatch with release atomics instead of __threadfence (no L1 invalidation): streamed-output + shared-frame tests, e2e breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m pytest tests/test_streamed_output_gpu.py tests/test_shared_host_frame_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4 | tee gpurun_out/r2g_stream_e2e.txt
