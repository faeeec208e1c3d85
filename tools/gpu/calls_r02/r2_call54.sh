#!/bin/bash
# bench at N GPUs (argument), e2e through the shared host frame
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > $O/r2R_bench_n$N.json 2> $O/r2R_bench_n$N.err; echo "bench rc $?"
tail -5 $O/r2R_bench_n$N.err
python - $N <<'P'
import json,sys
n=sys.argv[1]
d=json.loads(open('gpurun_out/r2R_bench_n%s.json'%n).read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], d.get('frame_check'))
e=d['e2e']; print('e2e', e['value'], e['ms_per_frame'], e.get('output'), e.get('copied_output'), e.get('lazy_ids'))
for k,w in d.get('workloads',{}).items():
    e=w.get('e2e',{}); print(k, w['ms_per_step'], w['value'], w.get('frame_check'), 'e2e', e.get('value'), e.get('ms_per_frame'), e.get('copied_output'))
P
