#!/bin/bash
# bench.py at N = 2 (torchrun, NCCL): frame check, config 4 sub-record, scene replication record
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r2z_bench_n2.json 2> $O/r2z_bench_n2.err ) 2>&1 | tail -3
tail -5 $O/r2z_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','frame_check')}); print(d['e2e']['value'], d['e2e']['ms_per_frame'])
print(json.dumps(d.get('scene_paths'),indent=1))
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v['ms_per_iteration'], v.get('frame_check'))
P
