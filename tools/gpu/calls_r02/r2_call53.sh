#!/bin/bash
# shared host frame (b200_stream_target, SceneHost::shareFrame): its 2-process tests, then the whole GPU suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_shared_host_frame_gpu.py tests/test_streamed_output_gpu.py -m gpu -q > $O/r2Q_shared_tests.log 2>&1; echo "shared tests rc $?" >> $O/r2Q_shared_tests.log
tail -40 $O/r2Q_shared_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2Q_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2Q_pytest.log
tail -3 $O/r2Q_pytest.log
