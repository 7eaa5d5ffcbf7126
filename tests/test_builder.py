"""Stereo local-map builder.  No reference code exists for this step, so the checker
is the numpy twin oracle/builder_ref.py (parity unpinned, see its header).
CPU: the twin recovers the truth within the measurement noise and its information blocks equal the
generator's; the CUDA kernel's per-landmark math (csrc/builder_math.h compiled for the host) equals the
twin's.  GPU: the CUDA builder equals the twin, and its maps drive the merge tree."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from linearsfm_b200 import builder, synth  # noqa: E402
from util import rel_err  # noqa: E402


@pytest.fixture(scope="module")
def obs():
    return builder.make_stereo_observations(6, feats_per_frame=40, seed=5)


def test_twin_recovers_truth_and_information(obs):
    import builder_ref as br
    pairs, cam, truth = obs
    for p, tr in zip(pairs, truth):
        lm, it = br.build_localmap(p, cam)
        assert it < 30                                                   # converged (linearly: noisy residuals), not cut off
        est = lm.stVal[:6]
        assert np.max(np.abs(est[:3] - tr["t_rel"])) < 0.05             # within the noise of ~40 landmarks
        assert np.max(np.abs(est[3:] - tr["a_rel"])) < 0.01
        # the information matrix is SPD and the Schur complement on the pose is well conditioned
        Vi = np.linalg.inv(lm.V)
        S = lm.U[0] - np.einsum("tij,tjk,tlk->il", lm.W, Vi, lm.W)
        assert np.all(np.linalg.eigvalsh(0.5 * (S + S.T)) > 0)
        assert np.array_equal(lm.feature, np.arange(lm.n)) and lm.m == 1 and lm.nU == 1


@pytest.fixture(scope="module")
def hostmath(tmp_path_factory):
    d = tmp_path_factory.mktemp("bm")
    so = str(d / "libbm.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "linearsfm_b200", "csrc"),
                           os.path.join(ROOT, "tests", "helpers", "builder_math_host.cpp"), "-o", so])
    L = C.CDLL(so)
    L.bm_feature_blocks.restype = C.c_double
    L.bm_inv3.restype = C.c_double
    return L


def test_kernel_math_equals_twin(hostmath, obs):
    import builder_ref as br
    pairs, cam, _ = obs
    p = pairs[2]
    X = br.triangulate(cam, p.z0)
    V, W, Ub, gF, gP = br.linearise(cam, p.pose0, X, p.z0, p.z1)
    pd = C.POINTER(C.c_double)
    for f in (0, 7, p.z0.shape[0] - 1):
        v, w, u, gf, gp = (np.zeros(k) for k in (9, 18, 36, 3, 6))
        args = [np.ascontiguousarray(a, float) for a in (p.pose0, X[f], p.z0[f], p.z1[f])]
        hostmath.bm_feature_blocks(C.c_double(cam.f), C.c_double(cam.b), C.c_double(cam.cx), C.c_double(cam.cy),
                                   C.c_double(cam.sigma), *[a.ctypes.data_as(pd) for a in args],
                                   *[a.ctypes.data_as(pd) for a in (v, w, u, gf, gp)])
        assert rel_err(v.reshape(3, 3), V[f]) < 1e-12
        assert rel_err(w.reshape(6, 3), W[f]) < 1e-12
        assert rel_err(u.reshape(6, 6), Ub[f]) < 1e-12
        assert rel_err(gf, gF[f]) < 1e-9 and rel_err(gp, gP[f]) < 1e-9
        xt = np.zeros(3)
        hostmath.bm_triangulate(C.c_double(cam.f), C.c_double(cam.b), C.c_double(cam.cx), C.c_double(cam.cy),
                                args[2].ctypes.data_as(pd), xt.ctypes.data_as(pd))
        assert rel_err(xt, X[f]) < 1e-14
    # 6x6 SPD solve and symmetric 3x3 inverse
    rng = np.random.default_rng(0)
    A = rng.normal(size=(6, 6)); S = A @ A.T + 6 * np.eye(6); e = rng.normal(size=6)
    Sx, ex = S.copy().reshape(-1), e.copy()
    assert hostmath.bm_solve6(Sx.ctypes.data_as(pd), ex.ctypes.data_as(pd)) == 1
    assert rel_err(ex, np.linalg.solve(S, e)) < 1e-12
    B = rng.normal(size=(3, 3)); Vm = B @ B.T + np.eye(3); o = np.zeros(9)
    hostmath.bm_inv3(np.ascontiguousarray(Vm).ctypes.data_as(pd), o.ctypes.data_as(pd))
    assert rel_err(o.reshape(3, 3), np.linalg.inv(Vm)) < 1e-12


@pytest.mark.gpu
def test_cuda_builder_equals_twin(gpu, obs):
    import builder_ref as br
    from linearsfm_b200.localmap import maps_equal_int
    pairs, cam, _ = obs
    got, iters = builder.build_localmaps_stereo(pairs, cam, max_iters=30, tol=1e-10)
    for p, g, it in zip(pairs, got, iters):
        ref, it_ref = br.build_localmap(p, cam, max_iters=30, tol=1e-10)
        assert not maps_equal_int(g, ref)
        assert abs(int(it) - it_ref) <= 2               # accept / reject and the stop test sit at rounding level near convergence
        assert rel_err(g.stVal, ref.stVal) < 1e-9
        for name in ("U", "W", "V"):
            assert rel_err(getattr(g, name), getattr(ref, name)) < 1e-8, name


@pytest.mark.gpu
def test_built_maps_drive_the_merge_tree(gpu, oracle):
    # raw observations -> CUDA builder -> CUDA merge tree, against builder twin -> reference tree
    import builder_ref as br
    pairs, cam, truth = builder.make_stereo_observations(16, feats_per_frame=48, seed=9)
    got_maps, _ = builder.build_localmaps_stereo(pairs, cam)
    ref_maps = [br.build_localmap(p, cam)[0] for p in pairs]
    got = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(got_maps)
    ref, _, _ = oracle.run_tree_stereo(ref_maps)
    assert np.array_equal(got.stno, ref.stno)
    assert rel_err(got.stVal, ref.stVal) < 1e-6
