#!/bin/bash
# streamed output (option 12): its tests, the whole GPU suite, bench at N = 1 (e2e streamed / copied / lazy ids)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_streamed_output_gpu.py -m gpu -q > $O/r2P_stream_tests.log 2>&1; echo "stream tests rc $?" >> $O/r2P_stream_tests.log
tail -25 $O/r2P_stream_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2P_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2P_pytest.log
tail -3 $O/r2P_pytest.log
timeout 900 python bench.py > $O/r2P_bench_n1.json 2> $O/r2P_bench_n1.err; echo "bench rc $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2P_bench_n1.json'))
print('value', d['value'], 'ms', d['ms_per_step'])
print(json.dumps(d['e2e'], indent=1))
for k,w in d.get('workloads',{}).items(): print(k, w['ms_per_step'], json.dumps(w.get('e2e',{}).get('ms_per_frame')), json.dumps(w.get('e2e',{}).get('copied_output')), json.dumps(w.get('e2e',{}).get('output')))
P
