#!/bin/bash
# compute-sanitizer memcheck over the kernels added since the last one: streamed output (both kernel instances, k_stream_own_tiles via the shared frame in one process), centre / half-extent nodes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_streamed_output_gpu.py "tests/test_shared_host_frame_gpu.py::test_one_process_shared_frame_equals_its_own_buffers" -m gpu -q -x > $O/r2a_memcheck.log 2>&1; echo "rc $?" >> $O/r2a_memcheck.log
grep -v "^$" $O/r2a_memcheck.log | tail -8
