#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_scene_broadcast_gpu.py tests/test_peer_frame_gpu.py -m gpu -q -x > $O/r2y_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2y_pytest.log
tail -25 $O/r2y_pytest.log
