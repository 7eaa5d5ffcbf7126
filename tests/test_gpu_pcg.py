"""north_star: "block-Jacobi PCG as a cross-check" of the GPU Cholesky that replaces CHOLMOD
(LinearSFMImp.cpp:2380-2449).  The reduced camera system S x = E of a join is captured from the
solve operator (lsfm_solve_stereo -> lsfm_debug_last_solve), solved again by preconditioned conjugate
gradients on the device (csrc/pcg.cu) and the two pose solutions are compared at 1e-8 relative."""
import numpy as np
import pytest

from linearsfm_b200 import synth

pytestmark = pytest.mark.gpu


def _joint_system(oracle, maps):
    """The reference's own join of a small tree, captured at the CHOLMOD boundary: the block system the
    reference hands to the solver (ea, eb, U, W, V of the joint map)."""
    recs = [r for r in oracle.run_levels_stereo(maps) if r["level"] != "final"]
    last = recs[-1]
    return last["Et"][0], last["C"][0], last["J"][0]


@pytest.mark.parametrize("n,fpf", [(8, 24), (32, 48), (64, 64)])
def test_pcg_agrees_with_cholesky(gpu, oracle, n, fpf):
    maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=300 + n, max_depth=15.0, gate=True)
    Et, Cm, J = _joint_system(oracle, maps)
    # right-hand sides of the joint system: b = I_End xhat_End (+) I_Cur xhat_Cur, assembled here from the
    # blocks the way lmj_LinearLS_PF3DStereo does (LinearSFMImp.cpp:2658-2930)
    m, nf = J.m, J.n
    ea = np.zeros((m, 6)); eb = np.zeros((nf, 3))
    order = np.argsort(J.feature_ids(), kind="stable")
    sid = J.feature_ids()[order]
    for S_, off in ((Et, 0), (Cm, Et.m)):
        xp, xf = S_.poses(), S_.features()
        fi = order[np.searchsorted(sid, S_.feature_ids())]
        for b in range(S_.nU):
            i, j = S_.Ui[b], S_.Uj[b]
            ea[off + i] += S_.U[b] @ xp[j]
            if i != j:
                ea[off + j] += S_.U[b].T @ xp[i]
        np.add.at(ea, off + S_.photo, np.einsum("bij,bj->bi", S_.W, xf[S_.feature]))
        np.add.at(eb, fi[S_.feature], np.einsum("bij,bi->bj", S_.W, xp[S_.photo]))
        np.add.at(eb, fi, np.einsum("fij,fj->fi", S_.V, xf))
    imp = gpu.CLinearSFMImp()
    st = imp.lmj_solveLinearSFMStereo(eb.reshape(-1), ea.reshape(-1), J.U, J.W, J.V, J.Ui, J.Uj, J.photo, J.feature, m, nf)
    # the operator reproduces the reference's joint state
    assert np.max(np.abs(st - J.stVal)) <= 1e-7 * np.max(np.abs(J.stVal))
    dbg = gpu.debug_last_solve()
    x, iters, rr = gpu.pcg_block(dbg["rowptr"], dbg["colidx"], dbg["S"], dbg["E"])
    chol = st[: 6 * m]
    err = np.max(np.abs(x - chol)) / np.max(np.abs(chol))
    print(f"PCG m={m}: {iters} iterations, relative residual {rr:.2e}, |x_pcg - x_chol| / |x| = {err:.2e}")
    assert rr <= 1e-10
    assert err <= 1e-8
