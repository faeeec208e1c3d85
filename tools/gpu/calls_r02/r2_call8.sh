#!/bin/bash
# Round 2, GPU call 8: the group walk (tracegroup.cuh) — first run: smoke, parity suite, timing of build variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; echo "smoke rc $?" >> $O/r2h_smoke.log
timeout 600 python tools/gpu/gpu_lib_sweep.py share > $O/r2h_lib_sweep.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/r2h_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2h_pytest.log
tail -3 $O/r2h_smoke.log; cat $O/r2h_lib_sweep.log; tail -15 $O/r2h_pytest.log
