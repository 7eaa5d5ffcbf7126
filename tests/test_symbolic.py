"""Host unit test of the symbolic phase of the block multifrontal Cholesky (linearsfm_b200/csrc/chol_symbolic.cpp,
the replacement of cholmod_amd + cholmod_analyze_p, LinearSFMImp.cpp:2413-2440): tests/helpers/symbolic_check.cpp
is compiled against the product source with g++ (no GPU, no oracle) and checks, on random batches of joins -- the
small-join fast path (m <= 32), mixed batches, large chains with hubs (LSFM-ND) and with loop closures (LSFM-MD) --
that the fronts contain the true fill of a dense boolean elimination, that relative indices, slot map, level sets,
front offsets, flop count are consistent (properties P1-P8 in the helper's header)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "linearsfm_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sym") / "symbolic_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", CSRC, os.path.join(ROOT, "tests", "helpers", "symbolic_check.cpp"),
                           os.path.join(CSRC, "chol_symbolic.cpp"), "-o", exe, "-lpthread"])
    return exe


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_symbolic_fronts_contain_the_fill(checker, seed):
    r = subprocess.run([checker, str(seed), "80"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "violations 0" in r.stdout, r.stdout[-2000:] + r.stderr[-500:]
