#!/bin/bash
# bench.py at N = 1 with the sub-workloads and the scene_paths record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python bench.py --gpus 1 > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err ) 2>&1 | tail -3
tail -3 $O/r2z_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','parity','clocks')})
print(json.dumps(d.get('scene_paths'),indent=1))
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v['ms_per_iteration'])
P
