#!/bin/bash
# Round-end measurement pass on the GPU box: tests, bench line, ncu launch list, ncu --set full of the
# four streaming kernels at the root level.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/fp_tests.log 2>&1; tail -3 gpurun_out/fp_tests.log
timeout 400 python bench.py > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err; tail -c 600 gpurun_out/fp_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/fp_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fp_bench_under_ncu.log 2>&1
for spec in "k_tf_chunk:12" "k_schur_pipe:11" "k_join_w:11" "k_backsub:11"; do
    kn=${spec%%:*}; skip=${spec##*:}
    timeout 150 ncu --set full --clock-control none -k regex:$kn --launch-skip $skip --launch-count 1 \
        -o gpurun_out/fp_$kn -f python tools/prof_one.py 3499 1 > gpurun_out/fp_ncu_$kn.log 2>&1
done
ls -la gpurun_out
