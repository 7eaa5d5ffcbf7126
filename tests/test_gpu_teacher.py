"""Teacher-forced parity at BASELINE.json's sizes (north_star: integers bit-exact, final poses / features
within 1e-6 relative, each joined objective within 1e-8 relative).

The reference's merge tree (LinearSFMImp.cpp:1926-2099) is run level by level with the reference's OWN
operators (oracle.run_levels_stereo); at every level the CUDA Transform / Join+Solve / re-base get exactly
the ORACLE's inputs of that level and are compared with the oracle's outputs of that level.  Errors of
one level therefore never reach the next, and every operator is pinned at the sizes where its chunk /
overflow / bitmap paths differ from the small cases -- including the root join of the headline config
(m = 3499 poses, ~6.2 M W blocks) and the final re-base.

Tolerances (relative to the largest entry of the array, `util.rel_err`):
  Transform  : state, U, W, V  <= 1e-9     (pure congruence / rigid transform, no solve; U, W, V <= 1e-8 for
                               maps of more than 512 poses, see tol_tf)
  Join       : U, W, V         <= 1e-12    (copies and sums of two blocks)
               state           <= 1e-9; a join that misses 1e-9 must stay within 20 x the reference's OWN
                               sensitivity of THAT join (the reference's join re-run three times on inputs
                               whose U, W, V carry 1e-15 relative noise, largest deviation: a solve cannot be
                               reproduced more closely than the reference reproduces itself).  north_star's
                               1e-6 is thereby enforced on every join whose reference sensitivity is below
                               5e-8; on the
                               open 3499-frame chain of the bench scene the reference itself moves by 4e-6 ..
                               7e-6 at level 8 join 1, 4e-5 at level 10 and 1e-3 at the root under that
                               noise (tests/tools/ref_sensitivity_levels.py), so no implementation can meet
                               1e-6 there -- test_closed_scene_3499_end_to_end asserts it on a
                               well-conditioned scene of the same size instead.  The report counts the
                               joins above 1e-9 and above 1e-6 per level.
  objective  : <= 1e-8 relative (a sample of <= 64 joins per level + every join that needed the sensitivity
               rule, which applies to its objective in the same way).
"""
import copy
import json
import os

import numpy as np
import pytest

from linearsfm_b200 import synth
from linearsfm_b200.localmap import maps_equal_int
from util import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def objective_blockwise(E, C, J):
    """F = sum_{k in {End, Cur}} (x|_k - xhat_k)^T I_k (x|_k - xhat_k), evaluated block by block
    (residual form) for the joint solution J of End E and Cur C."""
    jp = J.poses()
    jf = J.features()
    order = np.argsort(J.feature_ids(), kind="stable")
    sorted_ids = J.feature_ids()[order]
    F = 0.0
    for S, off in ((E, 0), (C, E.m)):
        dp = jp[off:off + S.m] - S.poses()
        fi = order[np.searchsorted(sorted_ids, S.feature_ids())]
        df = jf[fi] - S.features()
        u = np.einsum("bi,bij,bj->b", dp[S.Ui], S.U, dp[S.Uj])
        F += float(np.sum(np.where(S.Ui != S.Uj, 2.0 * u, u)))
        F += 2.0 * float(np.sum(np.einsum("bi,bij,bj->b", dp[S.photo], S.W, df[S.feature])))
        F += float(np.sum(np.einsum("fi,fij,fj->f", df, S.V, df)))
    return F


def _cmp(got, ref, what, worst, tol_state, tol_info):
    bad = maps_equal_int(got, ref)
    assert not bad, f"{what}: integer fields differ: {bad}"
    e = {"state": rel_err(got.stVal, ref.stVal), "U": rel_err(got.U, ref.U), "W": rel_err(got.W, ref.W),
         "V": rel_err(got.V, ref.V)}
    for k, v in e.items():
        worst[k] = max(worst.get(k, 0.0), v)
    assert e["state"] <= tol_state, f"{what}: state rel err {e['state']:.3e} > {tol_state:.3e}"
    for k in ("U", "W", "V"):
        assert e[k] <= tol_info, f"{what}: {k} rel err {e[k]:.3e} > {tol_info:.3e}"


def tol_tf(m):
    """U, W, V tolerance of a Transform: 1e-9 up to 512 poses, 1e-8 above -- U'(pos,pos) is a sum over every
    block of the map (millions of terms at the top levels, measured 2.5e-9 at m = 1024); the state stays
    at 1e-9 at every size."""
    return 1e-9 if m <= 512 else 1e-8


def run_teacher_forced(gpu, oracle, maps, tag, max_pairs_objective=64):
    report = []
    rng = np.random.default_rng(1)
    for rec in oracle.run_levels_stereo(maps):
        L = rec["level"]
        row = {"level": L}
        if L != "final":
            E, C, Et, J = rec["E"], rec["C"], rec["Et"], rec["J"]
            row.update(pairs=len(J), m=J[0].m, nW=int(max(j.nW for j in J)))
            # Transform(End -> Cur.Ref)
            w = {}
            got = gpu.transform_stereo_batch(E, [c.Ref for c in C])
            for i, (g, r) in enumerate(zip(got, Et)):
                _cmp(g, r, f"{tag} level {L} transform {i}", w, 1e-9, tol_tf(r.m))
            row["transform"] = w
            del got
            # Join + solve on the ORACLE's transformed End maps; objective of every join
            gpu.stats_reset(objective=True)
            try:
                got = gpu.join_stereo_batch(Et, C)
                obj = gpu.stats()["objectives"]
            finally:
                gpu.stats_reset()
            w = {}
            nsens, nloose, worst_ratio = 0, 0, 0.0
            obj_tol = {}
            for i, (g, r) in enumerate(zip(got, J)):
                e = rel_err(g.stVal, r.stVal)
                tol = 1e-9
                if e > tol:
                    # the reference's OWN sensitivity of this join: its result on inputs whose information
                    # blocks (U, W, V of both maps) carry 1e-15 relative noise
                    e2, c2 = copy.deepcopy(Et[i]), copy.deepcopy(C[i])
                    for mm in (e2, c2):
                        for nm in ("U", "W", "V"):
                            a = getattr(mm, nm)
                            setattr(mm, nm, a * (1 + 1e-15 * rng.standard_normal(a.shape)))
                    Fr0 = objective_blockwise(Et[i], C[i], r)
                    j2 = oracle.join_stereo(e2, c2)
                    sens = rel_err(j2.stVal, r.stVal)
                    osens = abs(objective_blockwise(Et[i], C[i], j2) - Fr0) / max(Fr0, 1e-300)
                    for _ in range(2):
                        for mm in (e2, c2):
                            for nm in ("U", "W", "V"):
                                a = getattr(mm, nm)
                                setattr(mm, nm, a * (1 + 1e-15 * rng.standard_normal(a.shape)))
                        j2 = oracle.join_stereo(e2, c2)
                        sens = max(sens, rel_err(j2.stVal, r.stVal))
                        osens = max(osens, abs(objective_blockwise(Et[i], C[i], j2) - Fr0) / max(Fr0, 1e-300))
                    obj_tol[i] = max(1e-8, 20.0 * osens)
                    tol = max(1e-9, 20.0 * sens)
                    nsens += 1
                    nloose += e > 1e-6
                    worst_ratio = max(worst_ratio, e / max(sens, 1e-300))
                _cmp(g, r, f"{tag} level {L} join {i}", w, tol, 1e-12)
            w["joins_above_1e-9"] = nsens
            w["joins_above_1e-6"] = nloose
            w["worst_err_over_ref_self_sensitivity"] = worst_ratio
            assert len(obj) == len(J)
            wo = 0.0
            step = max(1, len(J) // max_pairs_objective)
            for i in sorted(set(range(0, len(J), step)) | set(obj_tol)):
                Fr = objective_blockwise(Et[i], C[i], J[i])
                eo = abs(obj[i] - Fr) / max(Fr, 1e-300)
                # 1e-8 (north_star); joins that needed the sensitivity rule above get the same rule here:
                # 20 x the move of the reference's own objective under 1e-15 input noise
                assert eo <= obj_tol.get(i, 1e-8), f"{tag} level {L} join {i}: objective rel err {eo:.3e}"
                wo = max(wo, eo)
            w["objective"] = wo
            row["join"] = w
            del got
        if rec["rb_in"]:
            w = {}
            got = gpu.transform_stereo_batch(rec["rb_in"], rec["rb_ref"])
            for i, (g, r) in enumerate(zip(got, rec["rb_out"])):
                _cmp(g, r, f"{tag} level {L} re-base {i}", w, 1e-9, tol_tf(r.m))
            row["rebase"] = w
            del got
        report.append(row)
        print(json.dumps(row), flush=True)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"teacher_forced_{tag}.json"), "w") as f:
            json.dump(report, f, indent=1)
    return report


def test_teacher_forced_88(gpu, oracle):
    run_teacher_forced(gpu, oracle, synth.make_stereo_scene(88, feats_per_frame=64, seed=5), "n88")


def test_teacher_forced_3499_headline(gpu, oracle):
    """The bench workload itself (synthetic NC3500 shape, 3499 maps, 128 landmarks / frame)."""
    rep = run_teacher_forced(gpu, oracle, synth.make_stereo_scene(3499, feats_per_frame=128), "n3499")
    assert rep[-1]["level"] == "final" and rep[-2]["m"] == 3499


def test_closed_scene_3499_end_to_end(gpu, oracle):
    """Well-conditioned headline-size scene (loop closures every 500 frames, outlier-gated landmarks,
    SURVEY App. E): the WHOLE tree, no teacher forcing -- north_star's 1e-6 bar on the final state
    asserted directly, integers bit-exact."""
    maps = synth.make_stereo_scene(3499, feats_per_frame=128, revisit=0.1, lap=500, max_depth=15.0, gate=True)
    ref, _, _ = oracle.run_tree_stereo(maps)
    t = gpu.Tree(maps)
    t.solve()
    got = t.download(0)
    t.close()
    bad = maps_equal_int(got, ref)
    assert not bad, bad
    e = rel_err(got.stVal, ref.stVal)
    print(f"closed scene 3499: m={got.m} n={got.n} nU={got.nU} nW={got.nW} state rel err {e:.3e} "
          f"U {rel_err(got.U, ref.U):.3e} W {rel_err(got.W, ref.W):.3e}")
    assert e <= 1e-6


def test_two_runs_bit_identical(gpu):
    """Deterministic reductions (det_accum.cuh): the same solve twice gives the same BITS."""
    maps = synth.make_stereo_scene(466, feats_per_frame=128)
    res = []
    for _ in range(2):
        t = gpu.Tree(maps)
        t.solve()
        res.append(t.download(0))
        t.close()
    a, b = res
    assert not maps_equal_int(a, b)
    for name in ("stVal", "U", "W", "V"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), f"{name} differs between two runs"


def test_two_runs_bit_identical_on_the_slow_paths(gpu):
    """Chunks that see too many distinct poses (loop closures; forced here for every chunk with more than 4) take
    the thread-per-block paths of the Schur and Transform kernels: their sums are exact fixed-point integer
    accumulations (pre-pass for the scales), so two solves still give the same bits.  Fresh process: the
    switch is read once."""
    import subprocess, sys
    e = dict(os.environ); e["LSFM_FORCE_OVERFLOW"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dev_determinism.py"), "300", "64", "80"],
                       env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bit-identical: True" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]
