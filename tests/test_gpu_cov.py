"""marginal covariances of selected poses / landmarks of a joined map (the reference frees the
final information matrix unused, LinearSFMImp.cpp:2081-2096 -- there is no reference code, the check is the
definition: the blocks must be the diagonal blocks of the dense inverse of [[U, W], [W^T, V]]).

lsfm_marginal_cov_stereo / lsfm_tree_marginal_cov send unit right-hand sides through the joint solver of the
merge tree (Schur complement + multifrontal Cholesky + back-substitution)."""
import numpy as np
import pytest

from linearsfm_b200 import synth

pytestmark = pytest.mark.gpu


def _check(blocks, sel, Sigma, m, tol):
    for b, blk in zip(sel, blocks):
        d, o = (6, 6 * b) if b < m else (3, 6 * m + 3 * (b - m))
        ref = Sigma[o:o + d, o:o + d]
        assert blk.shape == (d, d)
        assert np.max(np.abs(blk - ref)) <= tol * np.max(np.abs(ref)), (b, np.max(np.abs(blk - ref)), np.max(np.abs(ref)))
        # a covariance: symmetric (to rounding), positive definite
        assert np.max(np.abs(blk - blk.T)) <= 1e-9 * np.max(np.abs(blk))
        assert np.all(np.linalg.eigvalsh(0.5 * (blk + blk.T)) > 0.0)


@pytest.mark.parametrize("n,fpf", [(8, 24), (24, 40)])
def test_marginal_cov_matches_dense_inverse(gpu, n, fpf):
    maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=500 + n, max_depth=15.0, gate=True)
    tree = gpu.Tree(maps)
    tree.solve()
    root = tree.download(0)
    m, nf = root.m, root.n
    Sigma = np.linalg.inv(root.dense_information())
    sel = [0, m // 2, m - 1, m, m + nf // 3, m + nf - 1]
    on_device = tree.marginal_cov(sel)                 # result map resident in HBM
    _check(on_device, sel, Sigma, m, 1e-6)
    from_host = gpu.marginal_cov(root, sel)            # the same map through host arrays
    for a, b in zip(on_device, from_host):
        assert np.array_equal(a, b)
    # more columns than one batch holds (48): every pose of the map
    allp = list(range(m))
    _check(tree.marginal_cov(allp), allp, Sigma, m, 1e-6)


def test_marginal_cov_leaf_map(gpu):
    leaf = synth.make_stereo_scene(4, feats_per_frame=16, seed=77)[1]
    Sigma = np.linalg.inv(leaf.dense_information())
    sel = [0, leaf.m, leaf.m + leaf.n - 1]
    _check(gpu.marginal_cov(leaf, sel), sel, Sigma, leaf.m, 1e-8)


def test_marginal_cov_rejects_bad_index(gpu):
    leaf = synth.make_stereo_scene(2, feats_per_frame=8, seed=5)[0]
    with pytest.raises(Exception):
        gpu.marginal_cov(leaf, [leaf.m + leaf.n])
