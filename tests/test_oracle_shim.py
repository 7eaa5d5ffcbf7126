"""Pins the CHOLMOD-ABI shim (the only restated third-party arithmetic, oracle/cholmod_shim.c):
its solution of the captured reduced camera system is checked by residual and against scipy."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from linearsfm_b200 import synth


def test_shim_solution_against_scipy(oracle):
    maps = synth.make_stereo_scene(40, feats_per_frame=24, seed=11)
    oracle.capture_enable(True)
    fin, _, _ = oracle.run_tree_stereo(maps)
    cap = oracle.capture()
    oracle.capture_enable(False)
    n = cap["n"]
    assert n == 6 * fin.m
    Aup = sp.csc_matrix((cap["Sx"], cap["Si"], cap["Sp"]), shape=(n, n))
    A = Aup + sp.triu(Aup, 1).T
    x = cap["x"]; b = cap["b"]
    r = A @ x - b
    assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(b)
    x2 = spla.spsolve(sp.csc_matrix(A), b)
    assert np.max(np.abs(x - x2)) <= 1e-8 * np.max(np.abs(x2))
    # scalar permutation = block permutation x6 (LinearSFMImp.cpp:2425-2434)
    assert np.array_equal(cap["perm"].reshape(-1, 6)[:, 0] // 6, cap["bperm"])
    # the root system (m=40 > 32) really exercised the dissection ordering
    assert not np.array_equal(cap["bperm"], np.arange(cap["m"]))
