#!/bin/bash
# shared-memory carve-out of the ray kernels (percent of 228 KB): the driver's choice against explicit values (6 CTAs x 17.4 KB = 105 KB are needed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
echo "driver's choice"; timeout 300 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep "^libvar"
for c in 46 50 60 75 100; do echo "carve-out $c %"; SOLR_B200_CARVEOUT=$c timeout 300 python tools/gpu/gpu_lib_sweep.py share 2>&1 | grep "^libvar"; done
