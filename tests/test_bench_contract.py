"""bench.py prints exactly ONE JSON line on stdout (everything else goes to stderr) with the keys the
driver reads.  CPU part: the reference arm (the reference's own CPU code through the oracle) on a tiny
prefix of the workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(oracle):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--maps", "48",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "s" and d["higher_is_better"] is False
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the arm measures the stated config: nothing scaled from a prefix, timed solves are counted
    assert d["extrapolated"] is False and "48 local maps" in d["config"]["workload"]
    assert d["steps"] == 1 and d["warmup"] == 0 and "all 48 local maps" in d["cpu_baseline"]["sample"]
