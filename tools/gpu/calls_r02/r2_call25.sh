#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SOLR_B200_TIMING=1 timeout 900 python tools/gpu/gpu_treebuild.py 2>&1 | tail -40
