#!/bin/bash
# what the streamed frames pay for: everything / counting without host writes / the counting kernels with nothing to count
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for d in 0 1 2; do echo "SOLR_B200_STREAM_DEBUG=$d"; SOLR_B200_STREAM_DEBUG=$d timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4 | grep streamed; done
