"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle (unmodified reference
TU + CHOLMOD shim) on the same seeded inputs.  Bars (BASELINE.json north_star): integer arrays
bit-exact; final poses/features <= 1e-6 relative; per-stage FP arrays <= 1e-9 relative here."""
import numpy as np
import pytest

from linearsfm_b200 import synth
from util import assert_maps_match, rel_err, state_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene16():
    return synth.make_stereo_scene(16, feats_per_frame=24, seed=7)


def test_transform_leaf(gpu, oracle, scene16):
    # leaf map: m=1, posID=0 absorbs every block (SURVEY Appendix E "Leaf Transform")
    for k in (0, 5):
        ref = oracle.transform_stereo(scene16[k], scene16[k + 1].Ref)
        got = gpu.transform_stereo_batch([scene16[k]], [scene16[k + 1].Ref])[0]
        assert_maps_match(got, ref, what=f"leaf transform {k}")


def test_transform_noop(gpu, oracle, scene16):
    got = gpu.transform_stereo_batch([scene16[3]], [scene16[3].Ref])[0]
    assert_maps_match(got, scene16[3], tol_state=0, tol_info=0, what="no-op transform")


def test_join_leaf_pair(gpu, oracle, scene16):
    e = oracle.transform_stereo(scene16[0], scene16[1].Ref)
    ref = oracle.join_stereo(e, scene16[1])
    got = gpu.join_stereo_batch([e], [scene16[1]])[0]
    assert_maps_match(got, ref, what="leaf join")


def test_transform_joint_map(gpu, oracle, scene16):
    # a level-1 map (m=2) re-based to its first frame: general routing of U/W blocks
    e = oracle.transform_stereo(scene16[2], scene16[3].Ref)
    j = oracle.join_stereo(e, scene16[3])
    assert j.Ref > j.FRef
    ref = oracle.transform_stereo(j, j.FRef)
    got = gpu.transform_stereo_batch([j], [j.FRef])[0]
    assert_maps_match(got, ref, what="re-base of a joint map")


def test_batched_transform_matches_single(gpu, oracle, scene16):
    maps = scene16[:6]
    refs = [scene16[i + 1].Ref for i in range(6)]
    got = gpu.transform_stereo_batch(maps, refs)
    for i in range(6):
        ref = oracle.transform_stereo(maps[i], refs[i])
        assert_maps_match(got[i], ref, what=f"batched transform {i}")


@pytest.mark.parametrize("n,fpf", [(2, 16), (3, 16), (5, 12), (8, 20), (16, 24), (37, 30)])
def test_tree_small(gpu, oracle, n, fpf):
    maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=100 + n)
    ref, _, _ = oracle.run_tree_stereo(maps)
    got = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(maps)
    assert_maps_match(got, ref, tol_state=1e-8, tol_info=1e-8, what=f"tree N={n}")
    assert state_rel_err(got, ref) <= 1e-6


def test_tree_88(gpu, oracle):
    maps = synth.make_stereo_scene(88, feats_per_frame=64, seed=88)
    ref, _, _ = oracle.run_tree_stereo(maps)
    t = gpu.Tree(maps)
    t.solve()
    got = t.download(0)
    t.close()
    assert_maps_match(got, ref, tol_state=1e-7, tol_info=1e-7, what="tree N=88")
    assert state_rel_err(got, ref) <= 1e-6


@pytest.mark.parametrize("env,size", [({"LSFM_FORCE_OVERFLOW": "1"}, ("37", "30")),
                                      ({"LSFM_SCHUR_DENSE": "1"}, ("37", "30")),
                                      ({"LSFM_SCHUR_DENSE": "1"}, ("300", "128")),
                                      ({"LSFM_SCHUR_V1": "1"}, ("37", "30")),
                                      ({"LSFM_FORCE_GLOBAL_PANEL": "1"}, ("88", "64")),
                                      ({"LSFM_BIGFRONT_MIN_FS": "96"}, ("300", "64")),
                                      ({"LSFM_BIGFRONT_MIN_FS": "96"}, ("480", "48", "closed80")),
                                      ({"LSFM_NO_BIGFRONT": "1"}, ("480", "48", "closed80"))])
def test_tree_alternative_paths(gpu, oracle, env, size):
    # LSFM_FORCE_OVERFLOW: chunks that see "too many" distinct poses (forced: > 4) take the thread-per-block
    # paths of the Transform / pattern / Schur kernels.  LSFM_SCHUR_DENSE: the DMMA Schur kernel
    # (schur_dense.cuh; 300 maps reach the 512-thread instantiation with 12 dense poses per chunk).
    # LSFM_SCHUR_V1: the first, atomics-only Schur kernel.  LSFM_FORCE_GLOBAL_PANEL: the Cholesky fronts'
    # global-memory panel (the path fronts taller than ~2100 rows take).  LSFM_BIGFRONT_MIN_FS=96: fronts of
    # 16 pose blocks or more take the multi-CTA path (chol_big.cuh: extend-add, panels, DMMA trailing update,
    # split back-solve) -- on the open chain (LSFM-ND fronts, one panel) and on a scene with loop closures
    # every 80 frames (LSFM-MD supernodes of several panels); LSFM_NO_BIGFRONT: the same scene with one CTA
    # per front.  Read once per process.
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "tools", "check_tree.py"), *size],
                       env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_solve_operator_and_pattern(gpu, oracle, scene16):
    # the narrowest pure-array operator (LinearSFMImp.h:209) + integer parity of the S pattern
    e = oracle.transform_stereo(scene16[0], scene16[1].Ref)
    j1 = oracle.join_stereo(e, scene16[1])
    e2 = oracle.transform_stereo(scene16[2], scene16[3].Ref)
    j2 = oracle.join_stereo(e2, scene16[3])
    j2b = oracle.transform_stereo(j2, j2.FRef)
    E = oracle.transform_stereo(j1, j2b.Ref)
    # build the joint system the reference would hand to the solver by running its join with
    # capture on, then feed the same (U,W,V,ea,eb) to both solve operators
    oracle.capture_enable(True)
    J = oracle.join_stereo(E, j2b)
    cap = oracle.capture()
    oracle.capture_enable(False)
    rng = np.random.default_rng(3)
    ea = rng.normal(size=6 * J.m); eb = rng.normal(size=3 * J.n)
    ref = oracle.solve_stereo(ea, eb, J.U, J.W, J.V, J.Ui, J.Uj, J.photo, J.feature, J.m, J.n)
    imp = gpu.CLinearSFMImp()
    got = imp.lmj_solveLinearSFMStereo(eb, ea, J.U, J.W, J.V, J.Ui, J.Uj, J.photo, J.feature, J.m, J.n)
    assert rel_err(got, ref) <= 1e-9
    dbg = gpu.debug_last_solve()
    # block pattern == the reference's Ap/Aii (upper CSC, LinearSFMImp.cpp:2529-2549) transposed to CRS
    m = cap["m"]
    assert dbg["m"] == m
    cols = [[] for _ in range(m)]
    for r in range(m):
        for q in range(dbg["rowptr"][r], dbg["rowptr"][r + 1]):
            cols[dbg["colidx"][q]].append(r)
    Ap = np.cumsum([0] + [len(c) for c in cols]).astype(np.int32)
    Ai = np.array([r for c in cols for r in c], np.int32)
    assert np.array_equal(Ap, cap["Ap"]) and np.array_equal(Ai, cap["Ai"])
    # ordering: GPU-side host routine == the oracle shim's independent implementation
    assert np.array_equal(dbg["perm"], oracle.shim_order(cap["Ap"], cap["Ai"]))
