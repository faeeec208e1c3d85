#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_p64.so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 8 --launch-count 4 --csv --log-file $O/r2o_launches_pool.csv python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2o_prof.log 2>&1
grep -v "^==" $O/r2o_launches_pool.csv | cut -d, -f5,13- | tail -21
SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_p64.so timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 5 --launch-count 1 -o $O/r2o_pool_pass1 -f python tools/gpu/prof_staged.py 2 4 3 2 > $O/r2o_ncu.log 2>&1
tail -2 $O/r2o_ncu.log
