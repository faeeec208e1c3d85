#!/bin/bash
# does the shared-memory footprint by itself cost instruction fetches?  unit walk with / without 16 KB of unused shared memory per CTA; pooled stages at 4 / 5 CTAs per SM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for run in "d0 1" "d16 1" "d0 2" "p64c4 2" "p64c5 2"; do
  set -- $run
  SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_$1.so timeout 300 ncu --metrics $M --clock-control none --launch-skip 8 --launch-count 3 --csv --log-file $O/r2p_$1_$2.csv python tools/gpu/prof_staged.py $2 4 3 4 > $O/r2p_prof.log 2>&1
  echo "== $1 mode $2"; grep -v "^==" $O/r2p_$1_$2.csv | cut -d, -f5,13- | grep -v "Kernel Name" | tr -d '"' | awk -F, '{printf "%s %s %s | ", $1, $4, $6} NR%5==0 {print ""}'
done
