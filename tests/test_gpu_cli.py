"""Drop-in process boundary: our LinearSFM executable vs the reference's own CLI binary
(oracle/_ref/LinearSFM_ref = reference main + implementation TU + CHOLMOD shim) on the same
localmap_<i>.txt files: same stdout progress lines, output files equal to the printed precision."""
import os
import subprocess

import numpy as np
import pytest

from linearsfm_b200 import _lib, synth
from linearsfm_b200.localmap import read_localmap, write_localmap

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF_CLI = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "LinearSFM_ref")


def run_cli(exe, d, n, typ, tag):
    out = {k: os.path.join(d, f"{tag}_{k}.txt") for k in ("p", "f", "st")}
    r = subprocess.run([exe, "-path", d, "-num", str(n), "-type", typ, "-p", out["p"], "-f", out["f"],
                        "-st", out["st"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout, out


def load_table(path):
    return np.array([[float(x) for x in line.split()] for line in open(path) if line.strip()])


@pytest.mark.parametrize("typ,n", [("Stereo", 6), ("Monocular", 5)])
def test_cli_dropin(gpu, tmp_path, typ, n):
    if not os.path.exists(REF_CLI):
        pytest.skip("reference CLI binary not built")
    mono = typ == "Monocular"
    maps = synth.make_mono_scene(n, 16, seed=31) if mono else synth.make_stereo_scene(n, 12, seed=32)
    import tempfile
    d = tempfile.mkdtemp(prefix="lsfm", dir="/tmp")        # the reference builds paths in char[200] (Imp.cpp:121)
    for i, lm in enumerate(maps):
        write_localmap(os.path.join(d, f"localmap_{i + 1}.txt"), lm, mono=mono)
    so_ref, f_ref = run_cli(REF_CLI, d, n, typ, "ref")
    so_got, f_got = run_cli(_lib.CLI_PATH, d, n, typ, "got")
    prog = lambda s: [l for l in s.splitlines() if l.startswith(("Join Level", "Generate Level"))]
    assert prog(so_got) == prog(so_ref)
    assert "Total Used Time:" in so_got
    for k in ("p", "f", "st"):
        a, b = load_table(f_got[k]), load_table(f_ref[k])
        assert a.shape == b.shape, k
        assert np.array_equal(a[:, 0], b[:, 0]), k                    # ids / stno
        assert np.max(np.abs(a[:, 1:] - b[:, 1:])) <= 2e-6, k        # %lf prints 6 decimals


def test_cli_errors(gpu):
    r = subprocess.run([_lib.CLI_PATH, "-num", "3", "-type", "Stereo"], capture_output=True, text=True)
    assert r.returncode == 0 and "Please Input Right File Path" in r.stdout
    r = subprocess.run([_lib.CLI_PATH, "-help"], capture_output=True, text=True)
    assert "Linear SFM Solution General Options" in r.stdout


def test_cli_map_output_can_be_joined_again(gpu, oracle, tmp_path):
    # -map <file> (extension): the joined map with its information matrix in the
    # reference's own localmap format -- the reference's reader gets back exactly what the tree computed
    from linearsfm_b200 import api
    from util import assert_maps_match
    import tempfile
    maps = synth.make_stereo_scene(6, 12, seed=32)
    d = tempfile.mkdtemp(prefix="lsfm", dir="/tmp")
    for i, lm in enumerate(maps):
        write_localmap(os.path.join(d, f"localmap_{i + 1}.txt"), lm)
    mp = os.path.join(d, "joined_map.txt")
    r = subprocess.run([_lib.CLI_PATH, "-path", d, "-num", "6", "-type", "Stereo", "-map", mp],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and os.path.exists(mp), r.stderr
    got = oracle.load_localmap_stereo(mp)                       # read by the REFERENCE's parser
    ref, _, _ = oracle.run_tree_stereo(maps)
    assert_maps_match(got, ref, tol_state=1e-8, tol_info=1e-8, what="-map output vs reference tree")


@pytest.mark.parametrize("typ,n", [("Stereo", 6), ("Monocular", 5)])
def test_cli_binary_cache(gpu, tmp_path, typ, n):
    # -cache <file> (extension): first run parses the text files and writes the cache, the second
    # run reads ONLY the cache (the text files are gone by then) and writes byte-identical outputs
    import tempfile
    mono = typ == "Monocular"
    maps = synth.make_mono_scene(n, 16, seed=31) if mono else synth.make_stereo_scene(n, 12, seed=32)
    d = tempfile.mkdtemp(prefix="lsfm", dir="/tmp")
    for i, lm in enumerate(maps):
        write_localmap(os.path.join(d, f"localmap_{i + 1}.txt"), lm, mono=mono)
    cache = os.path.join(d, "maps.lsfmcache")
    outs = []
    for run in range(2):
        o = {k: os.path.join(d, f"run{run}_{k}.txt") for k in ("p", "f", "st")}
        r = subprocess.run([_lib.CLI_PATH, "-path", d, "-num", str(n), "-type", typ, "-cache", cache,
                            "-p", o["p"], "-f", o["f"], "-st", o["st"]], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and os.path.exists(cache), r.stderr
        outs.append(o)
        for i in range(n):                                   # the second run must not need the text files
            p = os.path.join(d, f"localmap_{i + 1}.txt")
            if os.path.exists(p):
                os.remove(p)
    for k in ("p", "f", "st"):
        assert open(outs[0][k], "rb").read() == open(outs[1][k], "rb").read(), k


def test_cli_pose_covariances(gpu, tmp_path):
    # -cov <file> (extension): marginal 6x6 covariance of (at most 64 evenly spaced) poses of the
    # joined map; checked against the dense inverse of the information matrix the same run writes with -map
    import tempfile
    n = 6
    maps = synth.make_stereo_scene(n, 12, seed=41)
    d = tempfile.mkdtemp(prefix="lsfm", dir="/tmp")
    for i, lm in enumerate(maps):
        write_localmap(os.path.join(d, f"localmap_{i + 1}.txt"), lm)
    cov, mp = os.path.join(d, "cov.txt"), os.path.join(d, "joined.txt")
    r = subprocess.run([_lib.CLI_PATH, "-path", d, "-num", str(n), "-type", "Stereo", "-cov", cov, "-map", mp],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    joined = read_localmap(mp)
    Sigma = np.linalg.inv(joined.dense_information())
    rows = np.loadtxt(cov, ndmin=2)
    assert rows.shape == (joined.m, 37)
    ids = list(joined.pose_ids())
    for row in rows:
        p = ids.index(int(row[0]))
        ref = Sigma[6 * p:6 * p + 6, 6 * p:6 * p + 6]
        assert np.max(np.abs(row[1:].reshape(6, 6) - ref)) <= 1e-6 * np.max(np.abs(ref))
