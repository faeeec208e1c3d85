#!/bin/bash
# compute-sanitizer memcheck over the round's new kernels (GPU tree build, device animation, fused stages, scene replication arrays) on the small scenes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/r2K_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(gpu_built and config1) or (device_animation and config1) or (drivers_produce and config1) or host_container_animates" > $O/r2K_pytest.log 2>&1; echo "rc $?" >> $O/r2K_pytest.log
tail -3 $O/r2K_pytest.log; grep -c "Invalid\|ERROR SUMMARY" $O/r2K_memcheck.log; tail -3 $O/r2K_memcheck.log
