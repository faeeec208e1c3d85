#!/bin/bash
# last check of the round: the reference's own GPUKernel through B200Kernel (streams into its buffers), peer frame, bench headline with e2e
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m pytest tests/test_integration.py tests/test_peer_frame_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --no-sub --no-cpu-baseline > gpurun_out/r2h_bench_headline.json 2> gpurun_out/r2h_bench_headline.err; echo "bench rc $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2h_bench_headline.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['parity']); e=d['e2e']; print(e['value'], e['ms_per_frame'], e['copied_output'], e['lazy_ids']['ms_per_frame'])
P
