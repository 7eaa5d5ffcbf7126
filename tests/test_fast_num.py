"""The localmap reader's number parser (linearsfm_b200/csrc/fast_num.h; the reference reads the same files with
fscanf "%d" / "%lf", LinearSFMImp.cpp:3044-3132) must return exactly what strtod / strtol return: value bits and
end pointer, on random doubles in a dozen printf formats, random digit strings of 1-24 digits (the 16-19 digit
extended-precision path), exact double-precision ties, and edge cases (signs, bare points, dangling exponents,
hex floats, inf / nan, overflow, subnormals, leading white space)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_fast_num_equals_strtod(tmp_path, seed):
    exe = str(tmp_path / "fast_num_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "linearsfm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "helpers", "fast_num_check.cpp"), "-o", exe])
    r = subprocess.run([exe, str(seed), "1500000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-2000:]
