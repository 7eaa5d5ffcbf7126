"""Stereo local-map builder: ctypes mirror of lsfm_build_localmaps_stereo and a
seeded generator of the raw stereo observations it consumes.  The numpy twin used as the checker
lives under oracle/ (no reference code exists for this step: parity unpinned); nothing here imports it."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib, synth
from .localmap import LocalMap

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


class StereoPairC(C.Structure):
    _fields_ = [("Ref", C.c_int), ("pose_id", C.c_int), ("n", C.c_int), ("feat_id", _pi),
                ("z0", _pd), ("z1", _pd), ("pose0", _pd), ("X0", _pd)]


class StereoCamC(C.Structure):
    _fields_ = [("f", C.c_double), ("baseline", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("sigma", C.c_double)]


@dataclass
class StereoPair:
    """Raw input of one local map: the stereo measurements of the landmarks seen in two consecutive
    frames and an initial guess of the second frame's pose in the first frame."""
    Ref: int
    pose_id: int
    feat_id: np.ndarray      # int32 [n]
    z0: np.ndarray           # float64 [n, 3]  (uL, vL, uR) in the first frame
    z1: np.ndarray           # float64 [n, 3]  in the second frame
    pose0: np.ndarray        # float64 [6]
    X0: np.ndarray | None = None


def build_localmaps_stereo(pairs, cam: synth.StereoCam, max_iters: int = 30, tol: float = 1e-10):
    """CUDA path (C ABI).  Returns (list[LocalMap], iterations per map)."""
    from . import api
    L = _lib.lib()
    K = len(pairs)
    arr = (StereoPairC * K)()
    keep = []
    for k, p in enumerate(pairs):
        fid = np.ascontiguousarray(p.feat_id, np.int32)
        z0 = np.ascontiguousarray(p.z0, np.float64).reshape(-1)
        z1 = np.ascontiguousarray(p.z1, np.float64).reshape(-1)
        p0 = np.ascontiguousarray(p.pose0, np.float64).reshape(-1)
        keep += [fid, z0, z1, p0]
        arr[k].Ref, arr[k].pose_id, arr[k].n = int(p.Ref), int(p.pose_id), int(fid.shape[0])
        arr[k].feat_id = fid.ctypes.data_as(_pi)
        arr[k].z0 = z0.ctypes.data_as(_pd)
        arr[k].z1 = z1.ctypes.data_as(_pd)
        arr[k].pose0 = p0.ctypes.data_as(_pd)
        if p.X0 is not None:
            x0 = np.ascontiguousarray(p.X0, np.float64).reshape(-1)
            keep.append(x0)
            arr[k].X0 = x0.ctypes.data_as(_pd)
    cc = StereoCamC(cam.f, cam.b, cam.cx, cam.cy, cam.sigma)
    out = (_lib.LsfmMap * K)()
    iters = np.zeros(K, np.int32)
    api.check(L.lsfm_build_localmaps_stereo(arr, C.c_int(K), C.byref(cc), C.c_int(max_iters), C.c_double(tol),
                                            out, iters.ctypes.data_as(_pi)))
    return [api.from_c(out[k]) for k in range(K)], iters


def make_stereo_observations(num_maps: int, feats_per_frame: int = 128, seed: int = synth.SEED0,
                             min_life: int = 2, max_life: int = 6, cam: synth.StereoCam | None = None,
                             pose_noise: float = 0.02, min_disparity: float = 1.5):
    """Raw stereo measurements of a synthetic drive (same trajectory / landmark model as
    synth.make_stereo_scene): one StereoPair per pair of consecutive frames, pixel noise
    N(0, sigma^2), initial pose guess = truth + N(0, pose_noise^2).  Landmarks whose measured
    disparity falls below `min_disparity` pixels in either frame are not handed to the builder (the
    usual maximum-depth gate of a stereo front-end): their depth is unobservable, their V blocks are
    near-singular, and local maps that keep them make the joint systems of the map joining numerically
    indefinite -- for the reference's CPU code as well (non-positive pivots, NaN state at 466 maps).
    Returns (pairs, cam, truth) with truth = dict(t_rel, a_rel, X per pair)."""
    cam = cam or synth.StereoCam()
    rng = np.random.default_rng(np.random.PCG64(seed))
    N = int(num_maps)
    nf = N + 1
    p, ang, Rw = synth.make_trajectory(nf, rng)
    L_total = nf * feats_per_frame
    start = np.repeat(np.arange(nf), feats_per_frame)
    life = rng.integers(min_life, max_life + 1, L_total)
    end = np.minimum(start + life - 1, nf - 1)
    depth = rng.uniform(4.0, 30.0, L_total)
    th = rng.uniform(-np.deg2rad(28.0), np.deg2rad(28.0), L_total)
    tv = rng.uniform(-np.deg2rad(20.0), np.deg2rad(20.0), L_total)
    Xl = np.stack([depth, depth * np.tan(th), depth * np.tan(tv)], -1)
    Xw = np.einsum("nji,nj->ni", Rw[start], Xl) + p[start]
    gid = np.arange(1, L_total + 1, dtype=np.int32)
    pairs, truth = [], []
    for k in range(N):
        sel = np.flatnonzero((start <= k) & (end >= k + 1))
        Xk = np.einsum("ij,nj->ni", Rw[k], Xw[sel] - p[k])               # in frame k
        t_rel = Rw[k] @ (p[k + 1] - p[k])
        R_rel = Rw[k + 1] @ Rw[k].T
        a_rel = np.array(synth.ypr_from_rot(R_rel))
        Xc1 = np.einsum("ij,nj->ni", R_rel, Xk - t_rel)
        ok = (Xk[:, 0] > 1.0) & (Xc1[:, 0] > 1.0)                        # in front of both rigs
        sel, Xk, Xc1 = sel[ok], Xk[ok], Xc1[ok]
        z0 = cam.project(Xk) + rng.normal(0.0, cam.sigma, Xk.shape)
        z1 = cam.project(Xc1) + rng.normal(0.0, cam.sigma, Xk.shape)
        gate = (z0[:, 0] - z0[:, 2] >= min_disparity) & (z1[:, 0] - z1[:, 2] >= min_disparity)
        sel, Xk, z0, z1 = sel[gate], Xk[gate], z0[gate], z1[gate]
        pose0 = np.concatenate([t_rel, a_rel]) + rng.normal(0.0, pose_noise, 6) * np.array([1, 1, 1, 0.2, 0.2, 0.2])
        pairs.append(StereoPair(Ref=k + 1, pose_id=k + 2, feat_id=gid[sel], z0=z0, z1=z1, pose0=pose0))
        truth.append(dict(t_rel=t_rel, a_rel=a_rel, X=Xk))
    return pairs, cam, truth
