#!/bin/bash
# what the point-query tree (hits behind the ray origin, option 5) costs the walks: frames with and without it (timing only: the frames differ)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for o in 1 0; do echo "option 5 = $o"; SOLR_OPT5=$o timeout 300 python tools/gpu/gpu_variant_sweep.py 0 2>&1 | grep "^libvar"; done
