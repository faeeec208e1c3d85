#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "animat" > $O/r2D_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2D_pytest.log
tail -12 $O/r2D_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 > $O/r2D_bench_n1.json 2> $O/r2D_bench_n1.err; tail -2 $O/r2D_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2D_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}); print(json.dumps(d.get('scene_paths'),indent=1))
P
