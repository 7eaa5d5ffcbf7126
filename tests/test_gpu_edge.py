"""Edge cases of the join / tree against the oracle: the inputs the reference's own code paths
distinguish (LinearSFMImp.cpp:2565-2643: common / Cur-only features, first-index match of std::find;
1938-1947: odd map counts; leaves whose state already holds several poses)."""
import copy

import numpy as np
import pytest

from linearsfm_b200 import synth
from util import assert_maps_match, state_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    return synth.make_stereo_scene(12, feats_per_frame=20, seed=21)


def end_cur(oracle, scene, k=0):
    e = oracle.transform_stereo(scene[k], scene[k + 1].Ref)
    return e, copy.deepcopy(scene[k + 1])


def relabel(m, new_ids):
    m = copy.deepcopy(m)
    st = m.stno.copy()
    st[6 * m.m:] = np.repeat(np.asarray(new_ids, dtype=np.int32), 3)
    m.stno = st
    return m


def test_join_no_common_feature(gpu, oracle, scene):
    e, c = end_cur(oracle, scene)
    c = relabel(c, c.feature_ids() + 1_000_000)         # disjoint landmark ids: ncom = 0
    ref = oracle.join_stereo(e, c)
    assert ref.n == e.n + c.n
    got = gpu.join_stereo_batch([e], [c])[0]
    assert_maps_match(got, ref, what="join without common features")


def test_join_all_cur_features_common(gpu, oracle, scene):
    e, c = end_cur(oracle, scene)
    common = np.intersect1d(e.feature_ids(), c.feature_ids())
    assert len(common) >= 3
    # keep only the Cur features End also has (drop the others with their blocks)
    keep = np.isin(c.feature_ids(), common)
    idx = np.flatnonzero(keep)
    remap = -np.ones(c.n, dtype=np.int64); remap[idx] = np.arange(len(idx))
    wkeep = keep[c.feature]
    c2 = copy.deepcopy(c)
    c2.stno = np.concatenate([c.stno[:6 * c.m], np.repeat(c.feature_ids()[idx], 3)]).astype(np.int32)
    c2.stVal = np.concatenate([c.stVal[:6 * c.m], c.features()[idx].reshape(-1)])
    c2.n = len(idx)
    c2.V = c.V[idx]
    c2.W = c.W[wkeep]; c2.photo = c.photo[wkeep]; c2.feature = remap[c.feature[wkeep]].astype(np.int32)
    fb = np.full(c2.n, -1, dtype=np.int32)
    first = np.flatnonzero(np.r_[True, c2.feature[1:] != c2.feature[:-1]])
    fb[c2.feature[first]] = first
    c2.FBlock = fb
    ref = oracle.join_stereo(e, c2)
    assert ref.n == e.n                                  # nothing new comes from Cur
    got = gpu.join_stereo_batch([e], [c2])[0]
    assert_maps_match(got, ref, what="join with every Cur feature common")


def test_join_duplicate_id_first_match(gpu, oracle, scene):
    # two Cur features carry the same id as one End feature: std::find takes the FIRST (2581-2599),
    # the second stays a Cur-only feature
    e, c = end_cur(oracle, scene)
    common = np.intersect1d(e.feature_ids(), c.feature_ids())
    ids = c.feature_ids().copy()
    a = int(np.flatnonzero(ids == common[0])[0])
    b = int(np.flatnonzero(~np.isin(ids, common))[-1])   # a Cur-only feature gets the duplicate id
    assert a != b
    ids[b] = ids[a]
    c = relabel(c, ids)
    ref = oracle.join_stereo(e, c)
    got = gpu.join_stereo_batch([e], [c])[0]
    assert_maps_match(got, ref, what="join with a duplicated landmark id in Cur")


def test_tree_of_leaves_with_several_poses(gpu, oracle, scene):
    # leaves with m > 1: level-1 maps of the reference used as inputs of a new tree
    lvl1 = []
    for i in range(0, 12, 2):
        e = oracle.transform_stereo(scene[i], scene[i + 1].Ref)
        j = oracle.join_stereo(e, scene[i + 1])
        j.FRef = scene[i].Ref
        if j.Ref > j.FRef:
            j = oracle.transform_stereo(j, j.FRef)
        lvl1.append(j)
    assert all(m.m == 2 for m in lvl1)
    ref, _, _ = oracle.run_tree_stereo(lvl1)
    got = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(lvl1)
    assert_maps_match(got, ref, tol_state=1e-8, tol_info=1e-8, what="tree over m=2 leaves")
    assert state_rel_err(got, ref) <= 1e-6


def test_tree_single_map_is_identity(gpu, oracle, scene):
    # -num 1: the reference's level loop (1938) never runs and its final block reads an m_GMapS that
    # no join has filled (2039) -- undefined there (the oracle harness refuses it); here the single
    # map, already in its first frame, comes back unchanged
    got = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(scene[:1])
    assert_maps_match(got, scene[0], tol_state=0, tol_info=0, what="one local map")
