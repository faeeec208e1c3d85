#!/bin/bash
# Round 2, GPU call 12: the whole GPU suite (new: anaglyph on the staged driver, sample split, big-frame drop-in, host registration), smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=10 -k "not every_case" > $O/r2l_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2l_smoke.log 2>&1; echo "smoke rc $?" >> $O/r2l_smoke.log
tail -30 $O/r2l_pytest.log; tail -2 $O/r2l_smoke.log
