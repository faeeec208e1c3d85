#!/bin/bash
# streamed output counted per batch at the warp's convergent point: its tests, the e2e breakdown
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_shared_host_frame_gpu.py tests/test_streamed_output_gpu.py tests/test_peer_frame_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4 | tee $O/r2U_stream_e2e.txt
timeout 300 python tools/gpu/gpu_stream_e2e.py config4 2>&1 | tail -4 | tee -a $O/r2U_stream_e2e.txt
