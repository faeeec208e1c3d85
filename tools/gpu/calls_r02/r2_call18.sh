#!/bin/bash
# phased pooled walk: frames identical to the other drivers?  timing of the variants (queue depth, stuck threshold, pool size) against the staged driver
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" > $O/r2r_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2r_pytest.log
tail -5 $O/r2r_pytest.log
echo "== mode 1"; timeout 120 python tools/gpu/prof_staged.py 1 4 3 4 2>&1 | grep "^ms" | tail -2
echo "== mode 2"; timeout 120 python tools/gpu/prof_staged.py 2 4 3 4 2>&1 | grep "^ms" | tail -2
SOLR_MODE=2 timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2r_sweep.log 2>&1
cat $O/r2r_sweep.log
