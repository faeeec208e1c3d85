#!/bin/bash
# occupancy / shared-stack variants of the default build (csrc/libvar_*.so): 7 CTAs per SM (72 registers) for all / pass / primary kernels, 12- and 8-entry shared stacks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python tools/gpu/gpu_lib_sweep.py share 2>&1 > gpurun_out/r2S_variants.log 2>&1; grep "^libvar" gpurun_out/r2S_variants.log
