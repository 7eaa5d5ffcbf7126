"""Synthetic local-map generator (SURVEY.md section 8(d) "Concrete inputs").

The reference ships no datasets (DataForC/* hold Drive links only), so every test and bench input
is a seeded synthetic scene written in the reference's own local-map format.  Geometry follows the
reference's conventions (LinearSFMImp.cpp:132-143, 449-451): a pose is (t, alpha, beta, gamma),
R = Rx(gamma) Ry(beta) Rz(alpha) in the passive sense and a point maps as X_local = R (X - t).

Stereo scene ("NC-shape"): N+1 stereo frames on a smooth path; map k (1-based) is built from frames
(k, k+1): Ref = k, state = pose k+1 (6) + every landmark seen in both frames (3 each), all in the
frame of pose k; information = sum J' Sigma^-1 J of the 6 stereo measurements per landmark, i.e.
exactly the (U, W, V) a two-view bundle adjustment would hand to LinearSFM; the estimate is the
truth plus one Gauss-Newton-consistent draw of the measurement noise.
"""
from __future__ import annotations

import numpy as np

from .localmap import LocalMap

SEED0 = 20150328


def rot_ypr(a, b, g):
    """R(alpha, beta, gamma), row-major 3x3 (LinearSFMImp.cpp:132-143). Vectorised: [...,3,3]."""
    a, b, g = np.asarray(a, float), np.asarray(b, float), np.asarray(g, float)
    ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
    R = np.empty(a.shape + (3, 3))
    R[..., 0, 0] = cb * ca
    R[..., 0, 1] = cb * sa
    R[..., 0, 2] = -sb
    R[..., 1, 0] = sg * sb * ca - cg * sa
    R[..., 1, 1] = sg * sb * sa + cg * ca
    R[..., 1, 2] = sg * cb
    R[..., 2, 0] = cg * sb * ca + sg * sa
    R[..., 2, 1] = cg * sb * sa - sg * ca
    R[..., 2, 2] = cg * cb
    return R


def drot_ypr(a, b, g):
    """(dR/dalpha, dR/dbeta, dR/dgamma) of rot_ypr."""
    a, b, g = np.asarray(a, float), np.asarray(b, float), np.asarray(g, float)
    ca, sa, cb, sb, cg, sg = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(g), np.sin(g)
    z = np.zeros_like(ca)
    dA = np.stack([
        np.stack([-cb * sa, cb * ca, z], -1),
        np.stack([-sg * sb * sa - cg * ca, sg * sb * ca - cg * sa, z], -1),
        np.stack([-cg * sb * sa + sg * ca, cg * sb * ca + sg * sa, z], -1)], -2)
    dB = np.stack([
        np.stack([-sb * ca, -sb * sa, -cb], -1),
        np.stack([sg * cb * ca, sg * cb * sa, -sg * sb], -1),
        np.stack([cg * cb * ca, cg * cb * sa, -cg * sb], -1)], -2)
    dG = np.stack([
        np.stack([z, z, z], -1),
        np.stack([cg * sb * ca + sg * sa, cg * sb * sa - sg * ca, cg * cb], -1),
        np.stack([-sg * sb * ca + cg * sa, -sg * sb * sa - cg * ca, -sg * cb], -1)], -2)
    return dA, dB, dG


def ypr_from_rot(R):
    """Inverse of rot_ypr away from gimbal lock (LinearSFMImp.cpp:162-177)."""
    beta = np.arctan2(-R[..., 0, 2], np.sqrt(R[..., 0, 0] ** 2 + R[..., 0, 1] ** 2))
    cb = np.cos(beta)
    alpha = np.arctan2(R[..., 0, 1] / cb, R[..., 0, 0] / cb)
    gamma = np.arctan2(R[..., 1, 2] / cb, R[..., 2, 2] / cb)
    return alpha, beta, gamma


class StereoCam:
    """Pinhole stereo rig in the pose's local frame: x forward, y left, z up; the right camera sits
    at y = -baseline."""

    def __init__(self, f=400.0, baseline=0.12, cx=256.0, cy=192.0, sigma=0.5):
        self.f, self.b, self.cx, self.cy, self.sigma = f, baseline, cx, cy, sigma

    def project(self, Xc):
        x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
        return np.stack([self.cx - self.f * y / x, self.cy - self.f * z / x,
                         self.cx - self.f * (y + self.b) / x], -1)

    def jac(self, Xc):
        x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
        f, b = self.f, self.b
        J = np.zeros(Xc.shape[:-1] + (3, 3))
        J[..., 0, 0] = f * y / x ** 2
        J[..., 0, 1] = -f / x
        J[..., 1, 0] = f * z / x ** 2
        J[..., 1, 2] = -f / x
        J[..., 2, 0] = f * (y + b) / x ** 2
        J[..., 2, 1] = -f / x
        return J


def make_trajectory(nframes: int, rng: np.random.Generator):
    """World poses of the frames: position p[j] and angles (alpha, beta, gamma)[j]."""
    step = rng.uniform(0.3, 0.5, nframes)
    # smooth yaw rate: low-pass filtered noise, clipped to 5 deg / frame
    w = rng.normal(0.0, 1.0, nframes)
    k = np.exp(-np.arange(40) / 12.0)
    yaw_rate = np.convolve(w, k / k.sum(), mode="full")[:nframes]
    yaw_rate = np.clip(yaw_rate * np.deg2rad(6.0), -np.deg2rad(5.0), np.deg2rad(5.0))
    yaw = np.cumsum(yaw_rate)
    pitch = np.deg2rad(2.0) * np.sin(np.arange(nframes) * 0.11 + rng.uniform(0, 6.28))
    roll = np.deg2rad(1.5) * np.sin(np.arange(nframes) * 0.07 + rng.uniform(0, 6.28))
    ang = np.stack([yaw, pitch, roll], -1)
    R = rot_ypr(yaw, pitch, roll)
    fwd = R[:, 0, :]                           # local x axis expressed in world = first row of R
    p = np.zeros((nframes, 3))
    p[1:] = np.cumsum(fwd[:-1] * step[:-1, None], axis=0)
    p[:, 2] += 0.05 * np.sin(np.arange(nframes) * 0.05)
    return p, ang, R


def _stereo_block(cam, Rw, p, Xw, rows_l, rows_k, eps0_all, eps1_all, gate):
    """Two-view estimates and information blocks of the local maps k = 0..N-1 of one block: frames
    Rw[k], p[k] (N+1 of them), observation rows (rows_l = landmark, rows_k = map, sorted by map)."""
    N = Rw.shape[0] - 1
    sig = cam.sigma
    Rk, Rk1 = Rw[:N], Rw[1:N + 1]
    t_rel = np.einsum("kij,kj->ki", Rk, p[1:N + 1] - p[:N])
    R_rel = np.einsum("kij,klj->kil", Rk1, Rk)
    a_rel = np.stack(ypr_from_rot(R_rel), -1)
    w = 1.0 / sig ** 2

    for gate_pass in range(4 if gate else 1):
        n_of_map = np.bincount(rows_k, minlength=N)
        foff = np.concatenate([[0], np.cumsum(n_of_map)])
        if np.any(n_of_map == 0):
            raise ValueError("a local map has no features; increase feats_per_frame")
        eps0, eps1 = eps0_all, eps1_all
        # truth: relative pose of frame k+1 in frame k, landmark in frame k
        X0 = np.einsum("tij,tj->ti", Rk[rows_k], Xw[rows_l] - p[rows_k])      # in frame k

        def linearise(t_p, a_p, X):
            """Jacobians of the two stereo observations of every row at (pose, X)."""
            R1 = rot_ypr(a_p[:, 0], a_p[:, 1], a_p[:, 2])
            dA, dB, dG = drot_ypr(a_p[:, 0], a_p[:, 1], a_p[:, 2])
            R1r, dAr, dBr, dGr = R1[rows_k], dA[rows_k], dB[rows_k], dG[rows_k]
            d = X - t_p[rows_k]
            Xc1 = np.einsum("tij,tj->ti", R1r, d)
            J0 = cam.jac(X)                                   # obs in frame k: d z / d X
            Jp1 = cam.jac(Xc1)
            JX1 = np.einsum("tij,tjk->tik", Jp1, R1r)          # d z / d X
            Jang = np.stack([np.einsum("tij,tj->ti", dAr, d), np.einsum("tij,tj->ti", dBr, d),
                             np.einsum("tij,tj->ti", dGr, d)], -1)     # [T,3(xyz),3(angles)]
            JP1 = np.concatenate([-JX1, np.einsum("tij,tjk->tik", Jp1, Jang)], -1)   # [T,3,6]
            return J0, JX1, JP1, Xc1

        def assemble(J0, JX1, JP1):
            V = w * (np.einsum("tki,tkj->tij", J0, J0) + np.einsum("tki,tkj->tij", JX1, JX1))
            W = w * np.einsum("tki,tkj->tij", JP1, JX1)                  # [T,6,3]
            Ub = w * np.einsum("tki,tkj->tij", JP1, JP1)                 # [T,6,6]
            U = np.add.reduceat(Ub, foff[:-1], axis=0)
            return U, W, V

        # Gauss-Newton-consistent draw at the truth
        J0, JX1, JP1, _ = linearise(t_rel, a_rel, X0)
        U, W, V = assemble(J0, JX1, JP1)
        gF = w * (np.einsum("tki,tk->ti", J0, eps0) + np.einsum("tki,tk->ti", JX1, eps1))
        gP = np.add.reduceat(w * np.einsum("tki,tk->ti", JP1, eps1), foff[:-1], axis=0)
        Vi = np.linalg.inv(V)
        WVi = np.einsum("tij,tjk->tik", W, Vi)
        S = U - np.add.reduceat(np.einsum("tij,tkj->tik", WVi, W), foff[:-1], axis=0)
        e = gP - np.add.reduceat(np.einsum("tij,tj->ti", WVi, gF), foff[:-1], axis=0)
        dP = np.linalg.solve(S, e[..., None])[..., 0]
        dF = np.einsum("tij,tj->ti", Vi, gF - np.einsum("tji,tj->ti", W, dP[rows_k]))

        t_est = t_rel + dP[:, :3]
        a_est = a_rel + dP[:, 3:]
        X_est = X0 + dF
        if not gate:
            break
        _, _, _, Xc1e = linearise(t_est, a_est, X_est)
        bad = (X_est[:, 0] < 1.0) | (Xc1e[:, 0] < 1.0) | (np.abs(X_est[:, 0] - X0[:, 0]) > 0.5 * X0[:, 0])
        if not bad.any():
            break
        keep_rows = ~bad
        rows_l, rows_k = rows_l[keep_rows], rows_k[keep_rows]
        eps0_all, eps1_all = eps0_all[keep_rows], eps1_all[keep_rows]

    # information at the estimate = what the BA front-end would export
    J0, JX1, JP1, _ = linearise(t_est, a_est, X_est)
    U, W, V = assemble(J0, JX1, JP1)
    return dict(rows_l=rows_l, n_of_map=n_of_map, t_est=t_est, a_est=a_est, X_est=X_est, U=U, W=W, V=V)


def make_loop_trajectory(nframes: int, lap: int, rng: np.random.Generator):
    """Closed circuit driven lap after lap (`lap` frames per lap): frame j and frame j - lap see the
    same place from nearly the same pose (a few decimetres / tenths of a degree apart), which is what
    makes loop closures possible (the `--revisit` knob)."""
    j = np.arange(nframes)
    step = 0.4
    radius = lap * step / (2.0 * np.pi)
    lapno = j // lap
    phase = 2.0 * np.pi * (j % lap) / lap
    # every lap runs on its own slightly different line (radius / height / attitude offsets)
    lap_rng = np.random.default_rng(np.random.PCG64(int(rng.integers(1 << 31))))
    nl = int(lapno.max()) + 1
    dr = lap_rng.uniform(-0.4, 0.4, nl)[lapno]
    dz = lap_rng.uniform(-0.1, 0.1, nl)[lapno]
    rad = radius + dr + 0.15 * np.sin(3.0 * phase + lapno)
    p = np.stack([rad * np.sin(phase), rad * (1.0 - np.cos(phase)) - radius * 0.0, 0.05 * np.sin(5.0 * phase) + dz], -1)
    yaw = phase + np.deg2rad(1.0) * np.sin(7.0 * phase + 0.7 * lapno)
    # keep yaw continuous over the laps (the reference works on Euler angles; relative poses only)
    yaw = yaw + 2.0 * np.pi * lapno
    pitch = np.deg2rad(2.0) * np.sin(j * 0.11 + rng.uniform(0, 6.28))
    roll = np.deg2rad(1.5) * np.sin(j * 0.07 + rng.uniform(0, 6.28))
    ang = np.stack([yaw, pitch, roll], -1)
    R = rot_ypr(yaw, pitch, roll)
    return p, ang, R


def make_stereo_scene(num_maps: int, feats_per_frame: int = 128, seed: int = SEED0,
                      min_life: int = 2, max_life: int = 6, cam: StereoCam | None = None,
                      return_truth: bool = False, revisit: float = 0.0, lap: int = 500,
                      max_depth: float = 30.0, gate: bool = False, block_maps: int = 1024):
    """Generate `num_maps` consistent stereo local maps. Returns a list of LocalMap (views into
    flat arrays) and, optionally, the ground truth dict.

    revisit > 0: the rig drives laps of a closed circuit (`lap` frames per lap) and a
    fraction `revisit` of the landmarks is re-observed, with fresh measurement noise, when the rig
    passes the same place one lap later (and again on every further lap, each with probability
    `revisit`).  Those landmarks are shared between local maps that are far apart in the sequence:
    loop closures, which widen the band of the reduced camera system (larger Cholesky fronts) and
    anchor the chain.

    gate=True: outlier gating as a BA front-end would do it -- a landmark whose ESTIMATE in a local map
    lies closer than 1 m to either camera or whose estimated depth is off by more than 50 % is dropped
    from that map.  With max_depth=30 m (1.6 px disparity at 0.5 px noise) the linear noise draw below
    puts a few landmarks per thousand maps almost INTO a camera centre; their V blocks reach 1e14 and
    make the whole merge tree ill-conditioned (the reference's own result then moves by 1e-3 under
    1e-15 input noise).  The default (gate=False, max_depth=30) is kept as the
    round-1 bench workload; gate=True with max_depth=15 is the well-conditioned scene."""
    cam = cam or StereoCam()
    rng = np.random.default_rng(np.random.PCG64(seed))
    N = int(num_maps)
    nf = N + 1
    if revisit > 0.0:
        p, ang, Rw = make_loop_trajectory(nf, lap, rng)
    else:
        p, ang, Rw = make_trajectory(nf, rng)           # frame j (0-based) has pose id j+1

    # landmarks: spawned in frame j's frustum, tracked for `life` consecutive frames
    L_total = nf * feats_per_frame
    start = np.repeat(np.arange(nf), feats_per_frame)
    life = rng.integers(min_life, max_life + 1, L_total)
    end = np.minimum(start + life - 1, nf - 1)          # last frame (inclusive)
    depth = rng.uniform(4.0, max_depth, L_total)
    th = rng.uniform(-np.deg2rad(28.0), np.deg2rad(28.0), L_total)
    tv = rng.uniform(-np.deg2rad(20.0), np.deg2rad(20.0), L_total)
    Xl = np.stack([depth, depth * np.tan(th), depth * np.tan(tv)], -1)
    Xw = np.einsum("nji,nj->ni", Rw[start], Xl) + p[start]   # R^T Xl + p
    gid = np.arange(1, L_total + 1, dtype=np.int64)

    # (map k, landmark) rows: landmark belongs to maps k (0-based frame index of Ref) in [start, end-1]
    nmaps_of = np.maximum(end - start, 0)
    rows_l = np.repeat(np.arange(L_total), nmaps_of)
    offs = np.arange(nmaps_of.sum()) - np.repeat(np.cumsum(nmaps_of) - nmaps_of, nmaps_of)
    rows_k = start[rows_l] + offs
    keep = rows_k < N
    rows_l, rows_k = rows_l[keep], rows_k[keep]
    if revisit > 0.0:
        # re-observations one or more laps later: same landmark id, maps [start + q lap, end + q lap)
        ex_l, ex_k = [rows_l], [rows_k]
        alive = np.ones(L_total, bool)
        for q in range(1, nf // lap + 2):
            alive = alive & (rng.random(L_total) < revisit)
            sel = np.where(alive[rows_l])[0]
            kq = rows_k[sel] + q * lap
            ok = kq < N
            lq, kq = rows_l[sel][ok], kq[ok]
            if lq.size == 0:
                continue
            # the landmark must be in front of both cameras of the map (frames kq and kq + 1)
            xa = np.einsum("tj,tj->t", Rw[kq][:, 0, :], Xw[lq] - p[kq])
            xb = np.einsum("tj,tj->t", Rw[kq + 1][:, 0, :], Xw[lq] - p[kq + 1])
            front = (xa > 2.5) & (xb > 2.5) & (xa < 40.0) & (xb < 40.0)
            ex_l.append(lq[front]); ex_k.append(kq[front])
        rows_l, rows_k = np.concatenate(ex_l), np.concatenate(ex_k)
    order = np.lexsort((gid[rows_l], rows_k))
    rows_l, rows_k = rows_l[order], rows_k[order]
    sig = cam.sigma
    eps0_all = rng.normal(0.0, sig, (rows_k.shape[0], 3))
    eps1_all = rng.normal(0.0, sig, (rows_k.shape[0], 3))
    # The per-map work below is independent from map to map: it runs over blocks of `block_maps` maps so
    # that the 50k-map / multi-million-landmark scenes (BASELINE.json configs[4]) fit in host memory.
    if np.any(np.bincount(rows_k, minlength=N) == 0):
        raise ValueError("a local map has no features; increase feats_per_frame")
    bounds = np.searchsorted(rows_k, np.arange(0, N + block_maps, block_maps).clip(max=N))
    parts = []
    for bi in range(len(bounds) - 1):
        k0 = bi * block_maps
        k1 = min(N, k0 + block_maps)
        r0, r1 = bounds[bi], bounds[bi + 1]
        parts.append(_stereo_block(cam, Rw[k0:k1 + 1], p[k0:k1 + 1], Xw, rows_l[r0:r1], rows_k[r0:r1] - k0,
                                   eps0_all[r0:r1], eps1_all[r0:r1], gate))
    del eps0_all, eps1_all
    rows_l = np.concatenate([q["rows_l"] for q in parts])
    n_of_map = np.concatenate([q["n_of_map"] for q in parts])
    foff = np.concatenate([[0], np.cumsum(n_of_map)])
    t_est = np.concatenate([q["t_est"] for q in parts])
    a_est = np.concatenate([q["a_est"] for q in parts])
    X_est = np.concatenate([q["X_est"] for q in parts])
    U = np.concatenate([q["U"] for q in parts])
    W = np.concatenate([q["W"] for q in parts])
    V = np.concatenate([q["V"] for q in parts])
    del parts

    maps = []
    ids32 = gid.astype(np.int32)
    for k in range(N):
        s, e_ = foff[k], foff[k + 1]
        n = e_ - s
        stno = np.empty(6 + 3 * n, np.int32)
        stno[:6] = -(k + 2)
        stno[6:] = np.repeat(ids32[rows_l[s:e_]], 3)
        stVal = np.concatenate([t_est[k], a_est[k], X_est[s:e_].reshape(-1)])
        ar = np.arange(n, dtype=np.int32)
        maps.append(LocalMap(Ref=k + 1, stno=stno, stVal=stVal, m=1, n=int(n),
                             U=U[k:k + 1], Ui=np.zeros(1, np.int32), Uj=np.zeros(1, np.int32),
                             W=W[s:e_], photo=np.zeros(n, np.int32), feature=ar,
                             V=V[s:e_], FBlock=ar.copy()))
    if not return_truth:
        return maps
    # truth of the final map: everything in the frame of pose 1
    R0 = Rw[0]
    truth = {
        "pose_ids": np.arange(1, nf + 1),
        "pose_t": np.einsum("ij,kj->ki", R0, p - p[0]),
        "pose_R": np.einsum("kij,lj->kil", Rw, R0),
        "feat_ids": gid,
        "feat_X": np.einsum("ij,kj->ki", R0, Xw - p[0]),
    }
    return maps, truth


# ---------------------------------------------------------------------------------------------
# Monocular scene (SURVEY 8(d) "synthetic mono (RS-shape)", Appendix A.2 / C.3)
# ---------------------------------------------------------------------------------------------
def make_mono_scene(num_maps: int, feats_per_frame: int = 64, seed: int = SEED0 + 1,
                    min_life: int = 3, max_life: int = 6, sigma: float = 0.3, f: float = 400.0,
                    cx: float = 256.0, cy: float = 192.0, return_truth: bool = False,
                    style: str = "forward"):
    """`num_maps` monocular local maps.  Map k (1-based) is built from the three frames k, k+1, k+2:
    state = [pose k (= Ref, all zero), pose k+1 (= ScaP, translation component Fix pinned to Sign),
    pose k+2, landmarks], in the frame of pose k and in units where |t_{k+1}[Fix]| = 1; consecutive
    maps share two poses, as lmj_LinearLS_PF3DMono requires (LinearSFMImp.cpp:7383-7409).  The Ref
    and ScaP slots are adjacent (slots 0 and 1), which the reference's Transform silently relies on
    (SURVEY App. C.3).  Information = J' J / sigma^2 of the pixel observations (u, v) of every
    landmark in the frames of the map where it is tracked, with the gauge parameters (pose k, and
    component Fix of pose k+1) excluded, i.e. zero rows/columns."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    N = int(num_maps)
    nf = N + 2
    L_total = nf * feats_per_frame
    start = np.repeat(np.arange(nf), feats_per_frame)
    life = rng.integers(min_life, max_life + 1, L_total)
    end = np.minimum(start + life - 1, nf - 1)
    if style == "aerial":
        # RS-shape (aerial photogrammetry): the camera looks along +x at terrain ~100 units away and
        # flies sideways (+y) with a baseline/depth ratio of 0.2 -- strong parallax, 60-80 % overlap.
        B = 20.0
        p = np.zeros((nf, 3))
        p[:, 1] = B * np.arange(nf) + rng.normal(0, 0.5, nf)
        p[:, 0] = rng.normal(0, 1.0, nf)
        p[:, 2] = 3.0 * np.sin(np.arange(nf) * 0.3)
        ang = np.stack([np.deg2rad(3.0) * np.sin(np.arange(nf) * 0.21 + 1.0),
                        np.deg2rad(2.0) * np.sin(np.arange(nf) * 0.13), np.deg2rad(2.0) * np.sin(np.arange(nf) * 0.17)], -1)
        Rw = rot_ypr(ang[:, 0], ang[:, 1], ang[:, 2])
        Fix = 1
        Xw = np.stack([rng.uniform(80.0, 120.0, L_total),
                       p[start, 1] + 0.5 * (life - 1) * B + rng.uniform(-12.0, 12.0, L_total),
                       rng.uniform(-45.0, 45.0, L_total)], -1)
    else:
        p, ang, Rw = make_trajectory(nf, rng)
        Fix = 0
        depth = rng.uniform(4.0, 30.0, L_total)
        th = rng.uniform(-np.deg2rad(25.0), np.deg2rad(25.0), L_total)
        tv = rng.uniform(-np.deg2rad(18.0), np.deg2rad(18.0), L_total)
        Xl = np.stack([depth, depth * np.tan(th), depth * np.tan(tv)], -1)
        Xw = np.einsum("nji,nj->ni", Rw[start], Xl) + p[start]
    gid = np.arange(1, L_total + 1, dtype=np.int64)

    def proj(Xc):
        return np.stack([cx - f * Xc[..., 1] / Xc[..., 0], cy - f * Xc[..., 2] / Xc[..., 0]], -1)

    def jproj(Xc):
        x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
        J = np.zeros(Xc.shape[:-1] + (2, 3))
        J[..., 0, 0] = f * y / x ** 2; J[..., 0, 1] = -f / x
        J[..., 1, 0] = f * z / x ** 2; J[..., 1, 2] = -f / x
        return J

    maps = []
    truth_scale = np.zeros(N)
    for k in range(N):
        fr = [k, k + 1, k + 2]
        # landmarks tracked in at least two of the three frames
        vis = np.stack([(start <= j) & (end >= j) for j in fr], 0)        # [3, L]
        sel = np.where(vis.sum(0) >= 2)[0]
        n = sel.shape[0]
        if n < 8:
            raise ValueError("too few landmarks in a mono map; increase feats_per_frame")
        R0, p0 = Rw[k], p[k]
        rel_t = [R0 @ (p[j] - p0) for j in fr]
        rel_R = [Rw[j] @ R0.T for j in fr]
        s = 1.0 / abs(rel_t[1][Fix])
        Sign = 1 if rel_t[1][Fix] >= 0 else -1
        truth_scale[k] = s
        X = (Xw[sel] - p0) @ R0.T * s                                     # [n,3] in frame k, scaled
        poses = np.zeros((3, 6))
        for c in (1, 2):
            poses[c, :3] = rel_t[c] * s
            poses[c, 3:] = np.array(ypr_from_rot(rel_R[c]))
        # unknown vector: [pose1(6), pose2(6), X(3n)], gauge: pose1[Fix] fixed
        nun = 12 + 3 * n
        Jrows, rrows = [], []
        H = np.zeros((nun, nun))
        g = np.zeros(nun)
        for c in range(3):
            v = vis[c, sel]
            idx = np.where(v)[0]
            if idx.size == 0:
                continue
            if c == 0:
                Xc = X[idx]
                Jp = jproj(Xc)                                            # d z / d X
                eps = rng.normal(0, sigma, (idx.size, 2))
                for q, i in enumerate(idx):
                    sl = slice(12 + 3 * i, 15 + 3 * i)
                    H[sl, sl] += Jp[q].T @ Jp[q] / sigma ** 2
                    g[sl] += Jp[q].T @ eps[q] / sigma ** 2
            else:
                t, a = poses[c, :3], poses[c, 3:]
                R = rot_ypr(*a)
                dA, dB, dG = drot_ypr(*a)
                d = X[idx] - t
                Xc = d @ R.T
                Jp = jproj(Xc)
                JX = Jp @ R                                               # [q,2,3]
                Jang = np.stack([d @ dA.T, d @ dB.T, d @ dG.T], -1)       # [q,3,3]
                JP = np.concatenate([-JX, Jp @ Jang], -1)                 # [q,2,6]
                eps = rng.normal(0, sigma, (idx.size, 2))
                po = 6 * (c - 1)
                for q, i in enumerate(idx):
                    sl = slice(12 + 3 * i, 15 + 3 * i)
                    ps = slice(po, po + 6)
                    H[ps, ps] += JP[q].T @ JP[q] / sigma ** 2
                    H[ps, sl] += JP[q].T @ JX[q] / sigma ** 2
                    H[sl, ps] += JX[q].T @ JP[q] / sigma ** 2
                    H[sl, sl] += JX[q].T @ JX[q] / sigma ** 2
                    g[ps] += JP[q].T @ eps[q] / sigma ** 2
                    g[sl] += JX[q].T @ eps[q] / sigma ** 2
        # gauge: component Fix of pose 1 is not a variable
        H[Fix, :] = 0.0; H[:, Fix] = 0.0; g[Fix] = 0.0
        Hs = H.copy(); Hs[Fix, Fix] = 1.0
        delta = np.linalg.solve(Hs, g)
        poses[1] += delta[:6]; poses[2] += delta[6:12]
        Xe = X + delta[12:].reshape(n, 3)
        # information blocks (evaluated at the truth; adequate for a synthetic input)
        U = np.stack([H[0:6, 0:6], H[0:6, 6:12], H[6:12, 6:12]])
        Ui = np.array([1, 1, 2], np.int32); Uj = np.array([1, 2, 2], np.int32)
        Wb, ph, fe, FB = [], [], [], []
        for i in range(n):
            sl = slice(12 + 3 * i, 15 + 3 * i)
            first = -1
            for c in (1, 2):
                blk = H[6 * (c - 1):6 * c, sl]
                if vis[c, sel[i]]:
                    if first < 0:
                        first = len(Wb)
                    Wb.append(blk); ph.append(c); fe.append(i)
            FB.append(first)
        V = np.stack([H[12 + 3 * i:15 + 3 * i, 12 + 3 * i:15 + 3 * i] for i in range(n)])
        stno = np.concatenate([np.repeat([-(k + 1), -(k + 2), -(k + 3)], 6),
                               np.repeat(gid[sel].astype(np.int32), 3)]).astype(np.int32)
        stVal = np.concatenate([poses.reshape(-1), Xe.reshape(-1)])
        stVal[6 + Fix] = Sign
        lm = LocalMap(Ref=k + 1, stno=stno, stVal=stVal, m=3, n=int(n), U=U, Ui=Ui, Uj=Uj,
                      W=np.array(Wb), photo=np.array(ph, np.int32), feature=np.array(fe, np.int32),
                      V=V, FBlock=np.array(FB, np.int32), ScaP=k + 2, Fix=Fix, Sign=Sign,
                      FScaP=k + 2, FFix=Fix)
        maps.append(lm)
    if not return_truth:
        return maps
    R0 = Rw[0]
    s0 = truth_scale[0]
    truth = {"pose_ids": np.arange(1, nf + 1), "pose_t": (p - p[0]) @ R0.T * s0,
             "pose_R": np.einsum("kij,lj->kil", Rw, R0), "feat_ids": gid,
             "feat_X": (Xw - p[0]) @ R0.T * s0}
    return maps, truth
