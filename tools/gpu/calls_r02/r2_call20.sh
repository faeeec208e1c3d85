#!/bin/bash
# wavefront driver with the phased walk kernel: frames identical to the other drivers?  timing of the variants against the staged driver; per-kernel counters
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "drivers_produce" > $O/r2t_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2t_pytest.log
tail -5 $O/r2t_pytest.log
echo "== mode 1"; timeout 120 python tools/gpu/prof_staged.py 1 4 3 4 2>&1 | grep "^ms" | tail -2
echo "== mode 2"; timeout 120 python tools/gpu/prof_staged.py 2 4 3 4 2>&1 | grep "^ms" | tail -2
SOLR_MODE=2 timeout 900 python tools/gpu/gpu_lib_sweep.py share > $O/r2t_sweep.log 2>&1
cat $O/r2t_sweep.log
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
timeout 300 ncu --metrics $M --clock-control none --launch-skip 45 --launch-count 15 --csv --log-file $O/r2t_launches_wave.csv python tools/gpu/prof_staged.py 2 4 3 4 > $O/r2t_prof.log 2>&1
