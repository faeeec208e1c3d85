#!/bin/bash
# streamBatch out of line (the counting instance's registers): e2e breakdown, streamed-output tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4
SOLR_B200_STREAM_DEBUG=2 timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4 | grep "eager.*streamed"
timeout 600 python -m pytest tests/test_streamed_output_gpu.py tests/test_shared_host_frame_gpu.py -m gpu -q -x 2>&1 | tail -2
