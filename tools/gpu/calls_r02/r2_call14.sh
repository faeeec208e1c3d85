#!/bin/bash
# Round 2, GPU call 14: pooled stages (option 6 = 2) — frames identical to the other drivers?  timing against the staged driver, pool sizes 32 / 64 / 96.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" > $O/r2n_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2n_pytest.log
tail -5 $O/r2n_pytest.log
for v in p64 p32 p96; do
  for mode in 1 2; do
    echo "== $v mode $mode"; SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_$v.so timeout 120 python tools/gpu/prof_staged.py $mode 4 3 5 2>&1 | tail -3
  done
done > $O/r2n_timing.log 2>&1
cat $O/r2n_timing.log
