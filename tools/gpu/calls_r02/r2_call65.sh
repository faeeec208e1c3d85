#!/bin/bash
# final state of the round (centre / half-extent walk, streamed output, shared host frame): whole GPU suite, smoke, bench N = 1, ncu launch list of the bench command, full capture of one frame's four ray kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2b_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2b_pytest.log
tail -3 $O/r2b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py --gpus 1 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; tail -1 $O/r2b_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > $O/r2b_bench_under_ncu.log 2>&1
grep -c "k_stage" $O/r2b_launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage --launch-skip 4 --launch-count 4 -o $O/r2b_frame -f python tools/gpu/prof_staged.py 1 4 3 3 > $O/r2b_ncu.log 2>&1
tail -1 $O/r2b_ncu.log
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity','clocks')}); print(d['e2e']['value'], d['e2e']['ms_per_frame'], d['roofline']['frac'])
P
