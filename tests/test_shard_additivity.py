"""Basis of the planned feature-sharded top-level joins: a local map's
information is additive over its features, and the frame Transform is a congruence, so FEATURE SHARDS
OF A MAP ARE MAPS -- transforming the shards separately (the U blocks travel with one shard only) and
adding the results gives the transform of the whole map.  Checked here on the REFERENCE's own
Transform (oracle), i.e. the property belongs to the algorithm, not to our kernels.  CPU only."""
import copy

import numpy as np

from linearsfm_b200 import synth
from util import rel_err


def shard(lm, keep, with_u):
    s = copy.deepcopy(lm)
    idx = np.flatnonzero(keep)
    remap = -np.ones(lm.n, dtype=np.int64); remap[idx] = np.arange(len(idx))
    wkeep = keep[lm.feature]
    s.stno = np.concatenate([lm.stno[:6 * lm.m], np.repeat(lm.feature_ids()[idx], 3)]).astype(np.int32)
    s.stVal = np.concatenate([lm.stVal[:6 * lm.m], lm.features()[idx].reshape(-1)])
    s.n = len(idx)
    s.V = lm.V[idx]
    s.W = lm.W[wkeep]; s.photo = lm.photo[wkeep]; s.feature = remap[lm.feature[wkeep]].astype(np.int32)
    fb = np.full(s.n, -1, dtype=np.int32)
    first = np.flatnonzero(np.r_[True, s.feature[1:] != s.feature[:-1]])
    fb[s.feature[first]] = first
    s.FBlock = fb
    if not with_u:
        s.U = np.zeros_like(lm.U)           # same block list, no information: U travels with one shard
    return s


def test_transform_is_additive_over_feature_shards(oracle):
    maps = synth.make_stereo_scene(4, feats_per_frame=24, seed=33)
    # a level-2 map (4 poses, U / W fill-in from two levels of joins and re-bases), to be re-expressed
    # in the frame of one of its own poses
    M, _, _ = oracle.run_tree_stereo(maps)
    assert M.m == 4
    target = int(M.pose_ids()[2])
    keep_a = (M.feature_ids() % 2 == 0)
    Ma, Mb = shard(M, keep_a, True), shard(M, ~keep_a, False)
    T, Ta, Tb = (oracle.transform_stereo(x, target) for x in (M, Ma, Mb))
    # poses identical on every shard, landmarks partitioned
    assert rel_err(Ta.stVal[:6 * T.m], T.stVal[:6 * T.m]) < 1e-14
    assert rel_err(Tb.stVal[:6 * T.m], T.stVal[:6 * T.m]) < 1e-14
    ids = T.feature_ids()
    for S_, keep in ((Ta, keep_a), (Tb, ~keep_a)):
        pos = np.searchsorted(np.sort(ids), S_.feature_ids())
        order = np.argsort(ids)[pos]
        assert np.array_equal(ids[order], S_.feature_ids())
        assert rel_err(S_.features(), T.features()[order]) < 1e-13
        assert rel_err(S_.V, T.V[order]) < 1e-12
    # pose-pose information: the shard results ADD UP to the transform of the whole map
    assert np.array_equal(Ta.Ui, T.Ui) and np.array_equal(Tb.Uj, T.Uj)
    assert rel_err(Ta.U + Tb.U, T.U) < 1e-12
    # pose-landmark blocks: each shard holds exactly the blocks of its landmarks
    assert Ta.nW + Tb.nW == T.nW
    for S_ in (Ta, Tb):
        fid = S_.feature_ids()[S_.feature]
        key = {(int(p), int(f)): w for p, f, w in zip(T.photo, ids[T.feature], T.W)}
        for p, f, w in zip(S_.photo, fid, S_.W):
            assert rel_err(w, key[(int(p), int(f))]) < 1e-11


def test_reduced_system_is_additive_over_feature_shards(oracle):
    # S = U - sum_f W_f V_f^-1 W_f^T and E likewise: each shard contributes its landmarks' terms, the U
    # part comes from one shard -- what the planned sum-all-reduce of (S, E) relies on
    maps = synth.make_stereo_scene(4, feats_per_frame=24, seed=34)
    e = oracle.transform_stereo(maps[0], maps[1].Ref)
    J = oracle.join_stereo(e, maps[1])
    def reduced(lm):
        m = lm.m
        S = np.zeros((6 * m, 6 * m))
        for b in range(lm.nU):
            i, j = int(lm.Ui[b]), int(lm.Uj[b])
            S[6 * i:6 * i + 6, 6 * j:6 * j + 6] += lm.U[b]
            if i != j:
                S[6 * j:6 * j + 6, 6 * i:6 * i + 6] += lm.U[b].T
        Vi = np.linalg.inv(lm.V)
        for a in range(lm.nW):
            f = int(lm.feature[a])
            for b in np.flatnonzero(lm.feature == f):
                pa, pb = int(lm.photo[a]), int(lm.photo[b])
                S[6 * pa:6 * pa + 6, 6 * pb:6 * pb + 6] -= lm.W[a] @ Vi[f] @ lm.W[b].T
        return S
    keep = (J.feature_ids() % 3 == 0)
    Sa, Sb, S = reduced(shard(J, keep, True)), reduced(shard(J, ~keep, False)), reduced(J)
    assert rel_err(Sa + Sb, S) < 1e-12
