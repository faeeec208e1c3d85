#!/bin/bash
# after the clean-up: the whole GPU suite, smoke, timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2B_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2B_pytest.log
tail -4 $O/r2B_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for m in 1 2; do SOLR_MODE=$m timeout 300 python tools/gpu/gpu_share_sweep.py config2 2>&1 | grep share; done
