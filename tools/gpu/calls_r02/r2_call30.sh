#!/bin/bash
# bench.py at N = 8 (torchrun, NCCL)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=${1:-8}
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > $O/r2z_bench_n$N.json 2> $O/r2z_bench_n$N.err ) 2>&1 | tail -3
tail -3 $O/r2z_bench_n$N.err
python - $N <<'P'
import json,sys
d=json.loads(open('gpurun_out/r2z_bench_n%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','frame_check')}); print(d['e2e']['value'], d['e2e']['ms_per_frame'])
print(json.dumps(d.get('scene_paths'),indent=1))
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v['ms_per_iteration'], v.get('frame_check'), v['e2e']['ms_per_frame'])
P
