#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_integration.py tests/test_scene_broadcast_gpu.py -m gpu -q -x -k "animat or broadcast" > $O/r2N_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2N_pytest.log
tail -5 $O/r2N_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 > $O/r2N_bench_n1.json 2> $O/r2N_bench_n1.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2N_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}); print(json.dumps(d['scene_paths']['on_the_device']))
P
