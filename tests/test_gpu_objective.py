"""north_star: "each joined objective within 1e-8 relative".  The reference never evaluates the
objective of a join; the harness definition is the residual form
    F = sum_{k in {End, Cur}} (x|_k - xhat_k)^T I_k (x|_k - xhat_k)
with I_k the block information (U, W, V) of the two source maps, xhat_k their estimates and x the joint
solution.  Oracle side: the reference's join gives x, F is evaluated here in numpy from dense I_k;
CUDA side: k_objective (csrc/join.cu) evaluates it blockwise on the device for the library's own x."""
import numpy as np
import pytest

from linearsfm_b200 import synth

pytestmark = pytest.mark.gpu


def objective_ref(src_maps, joint):
    pose_at = {int(pid): i for i, pid in enumerate(joint.pose_ids())}
    feat_at = {int(fid): i for i, fid in enumerate(joint.feature_ids())}
    F = 0.0
    for S in src_maps:
        xs = np.concatenate([joint.poses()[[pose_at[int(p)] for p in S.pose_ids()]].reshape(-1),
                             joint.features()[[feat_at[int(f)] for f in S.feature_ids()]].reshape(-1)])
        d = xs - S.stVal
        F += float(d @ S.dense_information() @ d)
    return F


def test_join_objective_matches_reference(gpu, oracle):
    maps = synth.make_stereo_scene(8, feats_per_frame=20, seed=77)
    ends, curs, refs = [], [], []
    for i in range(0, 8, 2):
        e = oracle.transform_stereo(maps[i], maps[i + 1].Ref)
        ends.append(e); curs.append(maps[i + 1])
        refs.append(objective_ref([e, maps[i + 1]], oracle.join_stereo(e, maps[i + 1])))
    gpu.stats_reset(objective=True)
    try:
        gpu.join_stereo_batch(ends, curs)
        got = gpu.stats()["objectives"]
    finally:
        gpu.stats_reset()
    assert len(got) == len(refs)
    for g, r in zip(got, refs):
        assert r > 0
        assert abs(g - r) <= 1e-8 * r, (g, r)
