#!/bin/bash
# Round 2, GPU call 11: the new bench.py at N = 1 (headline + sub-records + parity child + reference arm).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > $O/r2k_bench_n1.json 2> $O/r2k_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/r2k_bench_ref.json 2> $O/r2k_bench_ref.err
tail -5 $O/r2k_bench_n1.err; cat $O/r2k_bench_n1.json | cut -c1-3000; tail -4 $O/r2k_bench_ref.err; cut -c1-400 $O/r2k_bench_ref.json
