#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python tools/gpu/gpu_lib_sweep.py 2>&1 | tail -2
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --metrics $M --clock-control none --launch-skip 8 --launch-count 3 --csv --log-file $O/r2w_launches.csv python tools/gpu/prof_staged.py 1 4 3 4 > $O/r2w_prof.log 2>&1
grep -v "^==" $O/r2w_launches.csv | cut -d, -f5,13- | tail -12
