"""TEST INFRASTRUCTURE ONLY -- numpy twin of the stereo local-map builder (csrc/builder.cu).

There is NO reference code for this step: LinearSFM starts from finished localmap_*.txt files
(DOC p.1), so this checker is a "port" of the repository's own specification and the builder's
parity is UNPINNED with respect to the reference.  It follows the same conventions
(LinearSFMImp.cpp:132-143 geometry; SURVEY Appendix A map layout) and the same Gauss-Newton schedule
as the CUDA kernel: linearise at the estimate, eliminate every landmark (S = sum Ub - W V^-1 W^T,
e = sum gP - W V^-1 gF), solve the 6x6 pose system, back-substitute (Levenberg-Marquardt damping with
accept / reject on the cost), stop when max|d pose| < tol.
Only tests/ and tools/ may import this module."""
from __future__ import annotations

import numpy as np

from linearsfm_b200 import synth
from linearsfm_b200.localmap import LocalMap


def linearise(cam, pose, X, z0, z1, with_cost=False):
    t, a = pose[:3], pose[3:]
    R1 = synth.rot_ypr(a[0], a[1], a[2])
    dA, dB, dG = synth.drot_ypr(a[0], a[1], a[2])
    d = X - t
    Xc1 = d @ R1.T
    J0 = cam.jac(X)
    Jp1 = cam.jac(Xc1)
    JX1 = Jp1 @ R1
    Jang = np.stack([d @ dA.T, d @ dB.T, d @ dG.T], -1)
    JP1 = np.concatenate([-JX1, np.einsum("tij,tjk->tik", Jp1, Jang)], -1)
    r0 = z0 - cam.project(X)
    r1 = z1 - cam.project(Xc1)
    w = 1.0 / cam.sigma ** 2
    V = w * (np.einsum("tki,tkj->tij", J0, J0) + np.einsum("tki,tkj->tij", JX1, JX1))
    W = w * np.einsum("tki,tkj->tij", JP1, JX1)
    Ub = w * np.einsum("tki,tkj->tij", JP1, JP1)
    gF = w * (np.einsum("tki,tk->ti", J0, r0) + np.einsum("tki,tk->ti", JX1, r1))
    gP = w * np.einsum("tki,tk->ti", JP1, r1)
    if with_cost:
        return V, W, Ub, gF, gP, float(w * (np.sum(r0 * r0) + np.sum(r1 * r1)))
    return V, W, Ub, gF, gP


def triangulate(cam, z):
    disp = np.maximum(z[:, 0] - z[:, 2], cam.f / 500.0)     # floor: a landmark 500 baselines away
    x = cam.f * cam.b / disp
    return np.stack([x, (cam.cx - z[:, 0]) * x / cam.f, (cam.cy - z[:, 1]) * x / cam.f], -1)


def build_localmap(pair, cam, max_iters=30, tol=1e-10):
    """Levenberg-Marquardt with the kernel's schedule: every pass linearises at the current estimate;
    if the cost went up since the last accepted point the step is undone and lambda grows tenfold,
    otherwise the point is accepted (lambda shrinks tenfold unless the previous pass was a rejection)
    and a damped step (V + lambda diag V per landmark, S + lambda diag S on the pose) is taken."""
    pose = np.array(pair.pose0, float).copy()
    X = np.array(pair.X0, float).copy() if pair.X0 is not None else triangulate(cam, pair.z0)
    lam, cost_prev, rejected = 1e-3, np.inf, False
    pose_b, X_b = pose.copy(), X.copy()
    it = 0
    while it < max_iters:
        it += 1
        V, W, Ub, gF, gP, cost = linearise(cam, pose, X, pair.z0, pair.z1, with_cost=True)
        if not (cost <= cost_prev * (1.0 + 1e-9)):
            pose, X = pose_b.copy(), X_b.copy()
            lam *= 10.0
            rejected = True
            continue
        pose_b, X_b, cost_prev = pose.copy(), X.copy(), cost
        if not rejected:
            lam = max(lam * 0.1, 1e-12)
        rejected = False
        idx = np.arange(3)
        Vd = V.copy()
        Vd[:, idx, idx] *= (1.0 + lam)
        Vi = np.linalg.inv(Vd)
        WVi = np.einsum("tij,tjk->tik", W, Vi)
        S = (Ub - np.einsum("tij,tkj->tik", WVi, W)).sum(0)
        S[np.arange(6), np.arange(6)] *= (1.0 + lam)
        e = (gP - np.einsum("tij,tj->ti", WVi, gF)).sum(0)
        dP = np.linalg.solve(S, e)
        X = X + np.einsum("tij,tj->ti", Vi, gF - np.einsum("tji,j->ti", W, dP))
        pose = pose + dP
        if np.max(np.abs(dP)) < tol:
            break
    V, W, Ub, _, _ = linearise(cam, pose, X, pair.z0, pair.z1)
    n = X.shape[0]
    ar = np.arange(n, dtype=np.int32)
    stno = np.concatenate([np.full(6, -pair.pose_id), np.repeat(np.asarray(pair.feat_id), 3)]).astype(np.int32)
    lm = LocalMap(Ref=pair.Ref, stno=stno, stVal=np.concatenate([pose, X.reshape(-1)]), m=1, n=n,
                  U=Ub.sum(0)[None], Ui=np.zeros(1, np.int32), Uj=np.zeros(1, np.int32), W=W,
                  photo=np.zeros(n, np.int32), feature=ar, V=V, FBlock=ar.copy())
    return lm, it
