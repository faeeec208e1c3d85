#!/bin/bash
# phased pooled walk: per-kernel counters and a source-level capture of pass 1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none --launch-skip 8 --launch-count 4 --csv --log-file $O/r2s_launches_phase.csv python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2s_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 8 --launch-count 2 -o $O/r2s_phase -f python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2s_ncu.log 2>&1
tail -2 $O/r2s_ncu.log
