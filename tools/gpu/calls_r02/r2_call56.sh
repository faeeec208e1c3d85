#!/bin/bash
# stage kernels in two instances (with / without the streamed-output hooks): GPU suite, e2e breakdown, bench at N = 1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2T_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2T_pytest.log
tail -3 $O/r2T_pytest.log
timeout 300 python tools/gpu/gpu_stream_e2e.py config2 2>&1 | tail -4 | tee $O/r2T_stream_e2e.txt
timeout 300 python tools/gpu/gpu_stream_e2e.py config4 2>&1 | tail -4 | tee -a $O/r2T_stream_e2e.txt
timeout 900 python bench.py > $O/r2T_bench_n1.json 2> $O/r2T_bench_n1.err; echo "bench rc $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2T_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline']['frac'])
e=d['e2e']; print('e2e', e['value'], e['ms_per_frame'], e.get('copied_output'), e.get('lazy_ids'))
for k,w in d.get('workloads',{}).items(): print(k, w['ms_per_step'], w.get('e2e',{}).get('ms_per_frame'), w.get('e2e',{}).get('copied_output',{}).get('ms_per_frame'))
P
