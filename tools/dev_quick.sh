#!/bin/bash
# Developer loop on the GPU box: the fast parity tests, then a short bench line (stage table on stdout).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS=${1:-"tests/test_gpu_parity.py tests/test_gpu_teacher.py::test_two_runs_bit_identical"}
timeout 300 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/quick_bench.json"))
    print("value %.2f ms  e2e %.2f ms" % (d["value"] * 1e3, d["e2e"]["value"] * 1e3))
    for k, v in sorted(d["stages"].items()):
        print("  %-18s %7.3f ms %4d launches %7.1f GB/s" % (k, v["ms"], v["launches"], v["GBps"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/quick_bench.err").read()[-1500:])
PY
