#!/bin/bash
# full capture of one STREAMED frame's four ray kernels (the counting instances k_stage_*<true>), for the stall reasons of what streaming costs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stage --launch-skip 64 --launch-count 4 -o $O/r2f_streamed_frame -f python tools/gpu/gpu_stream_e2e.py config2 > $O/r2f_ncu.log 2>&1
tail -3 $O/r2f_ncu.log; ls -la $O/r2f_streamed_frame.ncu-rep
