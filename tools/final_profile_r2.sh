#!/bin/bash
# Round-2 end-of-round measurement pass on the GPU box (outputs in gpurun_out/, copied to profiles/ by hand).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2f_tests.log 2>&1; tail -14 gpurun_out/r2f_tests.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.json
timeout 300 python bench.py --scene closed --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_closed.json 2> gpurun_out/r2f_bench_closed.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 300 gpurun_out/r2f_bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_bf_|k_front_" -c 3000 --csv \
    --log-file gpurun_out/r2f_chol_closed_launches.csv python tools/prof_one.py 3499 1 closed > /dev/null 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_schur_lock --launch-skip 10 --launch-count 1 \
    -o gpurun_out/r2f_lock_root -f python tools/prof_one.py 3499 1 > gpurun_out/r2f_lock_root.log 2>&1
tools/schur_levels.sh final
ls -la gpurun_out | tail -12
