#!/bin/bash
# Round 2, GPU call 13 (2 GPUs): bench.py at N = 2 through torchrun (headline + config4 sub-record, merged-frame checks).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > $O/r2m_bench_n2.json 2> $O/r2m_bench_n2.err
tail -5 $O/r2m_bench_n2.err; cut -c1-2500 $O/r2m_bench_n2.json
