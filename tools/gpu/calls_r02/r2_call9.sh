#!/bin/bash
# Round 2, GPU call 9: per-kernel times of the group walk vs the unit walk (launch lists), and one frame of the group walk under ncu --set full.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for v in 0_default a_unit; do
  SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_$v.so timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none --launch-skip 8 --launch-count 4 --csv --log-file $O/r2i_launches_$v.csv python tools/gpu/prof_staged.py 1 4 3 4 > $O/r2i_prof_$v.log 2>&1
done
SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_0_default.so timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 3 -o $O/r2i_group_frame -f python tools/gpu/prof_staged.py 1 4 3 2 > $O/r2i_ncu.log 2>&1
grep -v "^==" $O/r2i_launches_0_default.csv | cut -d, -f5,13- | tail -14; grep -v "^==" $O/r2i_launches_a_unit.csv | cut -d, -f5,13- | tail -14; tail -3 $O/r2i_ncu.log
