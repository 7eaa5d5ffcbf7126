#!/bin/bash
# Round-2 (second half) end-of-round measurement pass on the GPU box; $1 = tests | bench
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
    timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2g_tests.log 2>&1; tail -14 gpurun_out/r2g_tests.log | cut -c1-200
else
    timeout 400 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 600 gpurun_out/r2g_bench.json
    timeout 300 python bench.py --scene closed --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_closed.json 2> gpurun_out/r2g_bench_closed.err
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2g_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_bench_under_ncu.log 2>&1
    python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
fi
ls -la gpurun_out | tail -8
