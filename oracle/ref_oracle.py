"""ctypes binding of the CPU oracle (oracle/_ref/libref_linearsfm.so).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package (linearsfm_b200) never does.

The library is the UNMODIFIED reference TU (LinearSFMImp.cpp) + oracle/cholmod_shim.c, built by
oracle/Makefile (see oracle/ref_harness.cpp for the exact reference entry points each call runs).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linearsfm_b200.localmap import LocalMap  # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_linearsfm.so")
CLI_PATH = os.path.join(_HERE, "_ref", "LinearSFM_ref")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


class RefMap(C.Structure):
    _fields_ = [("Ref", C.c_int), ("FRef", C.c_int), ("r", C.c_int), ("m", C.c_int), ("n", C.c_int),
                ("nU", C.c_int), ("nW", C.c_int),
                ("ScaP", C.c_int), ("Fix", C.c_int), ("Sign", C.c_int), ("FScaP", C.c_int),
                ("FFix", C.c_int),
                ("stno", _pi), ("stVal", _pd), ("U", _pd), ("Ui", _pi), ("Uj", _pi),
                ("W", _pd), ("photo", _pi), ("feature", _pi), ("V", _pd), ("FBlock", _pi)]


class ShimCapture(C.Structure):
    _fields_ = [("m", C.c_int), ("Ap", _pi), ("Ai", _pi), ("bperm", _pi), ("n", C.c_int),
                ("Sp", _pi), ("Si", _pi), ("Sx", _pd), ("perm", _pi), ("b", _pd), ("x", _pd),
                ("n_amd", C.c_long), ("n_factorize", C.c_long), ("n_solve", C.c_long),
                ("t_amd", C.c_double), ("t_analyze", C.c_double), ("t_factorize", C.c_double),
                ("t_solve", C.c_double)]


def build(force: bool = False) -> bool:
    """Build oracle/_ref from the reference sources if they are present (this container)."""
    ref = os.environ.get("LSFM_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        return os.path.exists(LIB_PATH)
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, f"REF={ref}"], stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"oracle library missing: {LIB_PATH} (run `make -C oracle`)")
        _lib = C.CDLL(LIB_PATH)
        _lib.lsfm_shim_capture.restype = C.POINTER(ShimCapture)
        _lib.lsfm_shim_time.restype = C.c_double
    return _lib


def _keep(arr, dtype):
    return np.ascontiguousarray(arr, dtype=dtype)


def to_c(lm: LocalMap):
    """LocalMap -> (RefMap, keepalive list). Arrays are passed by pointer; harness deep-copies."""
    ka = dict(stno=_keep(lm.stno, np.int32), stVal=_keep(lm.stVal, np.float64),
              U=_keep(lm.U, np.float64), Ui=_keep(lm.Ui, np.int32), Uj=_keep(lm.Uj, np.int32),
              W=_keep(lm.W, np.float64), photo=_keep(lm.photo, np.int32),
              feature=_keep(lm.feature, np.int32), V=_keep(lm.V, np.float64),
              FBlock=_keep(lm.FBlock, np.int32))
    c = RefMap(Ref=lm.Ref, FRef=lm.FRef, r=lm.r, m=lm.m, n=lm.n, nU=lm.nU, nW=lm.nW,
               ScaP=lm.ScaP, Fix=lm.Fix, Sign=lm.Sign, FScaP=lm.FScaP, FFix=lm.FFix)
    for k, v in ka.items():
        setattr(c, k, v.ctypes.data_as(_pi if v.dtype == np.int32 else _pd))
    return c, ka


def from_c(c: RefMap, free: bool = True) -> LocalMap:
    def ai(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)

    def ad(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0, np.float64)

    r = 6 * c.m + 3 * c.n
    lm = LocalMap(Ref=c.Ref, FRef=c.FRef, stno=ai(c.stno, r), stVal=ad(c.stVal, r), m=c.m, n=c.n,
                  U=ad(c.U, 36 * c.nU), Ui=ai(c.Ui, c.nU), Uj=ai(c.Uj, c.nU),
                  W=ad(c.W, 18 * c.nW), photo=ai(c.photo, c.nW), feature=ai(c.feature, c.nW),
                  V=ad(c.V, 9 * c.n), FBlock=ai(c.FBlock, c.n),
                  ScaP=c.ScaP, Fix=c.Fix, Sign=c.Sign, FScaP=c.FScaP, FFix=c.FFix)
    if free:
        lib().ref_free_map(C.byref(c))
    return lm


def transform_stereo(lm: LocalMap, Ref: int) -> LocalMap:
    """CLinearSFMImp::lmj_Transform_PF3DStereo (LinearSFMImp.cpp:349)."""
    c, _ka = to_c(lm)
    out = RefMap()
    rc = lib().ref_transform_stereo(C.byref(c), C.c_int(Ref), C.byref(out))
    assert rc == 0
    return from_c(out)


def join_stereo(end: LocalMap, cur: LocalMap) -> LocalMap:
    """CLinearSFMImp::lmj_LinearLS_PF3DStereo (LinearSFMImp.cpp:2551)."""
    a, _k1 = to_c(end)
    b, _k2 = to_c(cur)
    out = RefMap()
    rc = lib().ref_join_stereo(C.byref(a), C.byref(b), C.byref(out))
    assert rc == 0
    return from_c(out)


def solve_stereo(ea, eb, U, W, V, Ui, Uj, photo, feature, m, n):
    """CLinearSFMImp::lmj_solveLinearSFMStereo (LinearSFMImp.cpp:2119). Returns stVal[6m+3n]."""
    ea = _keep(ea, np.float64).copy(); eb = _keep(eb, np.float64).copy()
    U = _keep(U, np.float64).copy(); W = _keep(W, np.float64).copy(); V = _keep(V, np.float64).copy()
    Ui = _keep(Ui, np.int32).copy(); Uj = _keep(Uj, np.int32).copy()
    photo = _keep(photo, np.int32).copy(); feature = _keep(feature, np.int32).copy()
    st = np.zeros(6 * m + 3 * n)
    p = lambda a: a.ctypes.data_as(_pi if a.dtype == np.int32 else _pd)
    lib().ref_solve_stereo(p(st), p(eb), p(ea), p(U), p(W), p(V), p(Ui), p(Uj), p(photo),
                           p(feature), C.c_int(m), C.c_int(n), C.c_int(Ui.shape[0]),
                           C.c_int(photo.shape[0]))
    return st


def _run_tree(fn, maps):
    arr = (RefMap * len(maps))()
    keep = []
    for i, lm in enumerate(maps):
        c, ka = to_c(lm)
        arr[i] = c
        keep.append(ka)
    out = RefMap()
    t_ref, t_wall = C.c_double(0), C.c_double(0)
    rc = fn(arr, C.c_int(len(maps)), C.byref(out), C.byref(t_ref), C.byref(t_wall))
    if rc != 0:
        raise RuntimeError(f"reference tree run failed rc={rc}")
    return from_c(out), t_ref.value, t_wall.value


def run_tree_stereo(maps):
    """CLinearSFMImp::lmj_PF3D_Divide_ConquerStereo (LinearSFMImp.cpp:1926).
    Returns (final LocalMap, reference's own clock() seconds, wall seconds)."""
    return _run_tree(lib().ref_run_tree_stereo, maps)


def transform_mono(lm: LocalMap, Ref: int, ScaP: int, Fix: int) -> LocalMap:
    c, _ka = to_c(lm)
    out = RefMap()
    rc = lib().ref_transform_mono(C.byref(c), C.c_int(Ref), C.c_int(ScaP), C.c_int(Fix),
                                  C.byref(out))
    assert rc == 0
    return from_c(out)


def join_mono(end: LocalMap, cur: LocalMap) -> LocalMap:
    a, _k1 = to_c(end)
    b, _k2 = to_c(cur)
    out = RefMap()
    rc = lib().ref_join_mono(C.byref(a), C.byref(b), C.byref(out))
    assert rc == 0
    return from_c(out)


def solve_mono(ea, eb, U, W, V, Ui, Uj, photo, feature, m, n, Ref, ScaP, Fix, Sign, FixBlk):
    """CLinearSFMImp::lmj_solveLinearSFMMono (LinearSFMImp.cpp:6756). Returns stVal[6m+3n]."""
    ea = _keep(ea, np.float64).copy(); eb = _keep(eb, np.float64).copy()
    U = _keep(U, np.float64).copy(); W = _keep(W, np.float64).copy(); V = _keep(V, np.float64).copy()
    Ui = _keep(Ui, np.int32).copy(); Uj = _keep(Uj, np.int32).copy()
    photo = _keep(photo, np.int32).copy(); feature = _keep(feature, np.int32).copy()
    st = np.zeros(6 * m + 3 * n)
    p = lambda a: a.ctypes.data_as(_pi if a.dtype == np.int32 else _pd)
    lib().ref_solve_mono(p(st), p(eb), p(ea), p(U), p(W), p(V), p(Ui), p(Uj), p(photo), p(feature),
                         C.c_int(m), C.c_int(n), C.c_int(Ui.shape[0]), C.c_int(photo.shape[0]),
                         C.c_int(Ref), C.c_int(ScaP), C.c_int(Fix), C.c_int(Sign), C.c_int(FixBlk))
    return st


def run_tree_mono(maps):
    return _run_tree(lib().ref_run_tree_mono, maps)


def load_localmap_stereo(path: str) -> LocalMap:
    out = RefMap()
    rc = lib().ref_load_localmap_stereo(path.encode(), C.byref(out))
    if rc != 0:
        raise FileNotFoundError(path)
    return from_c(out)


def save_outputs(lm: LocalMap, st=None, pose=None, feat=None):
    c, _ka = to_c(lm)
    lib().ref_save_outputs(C.byref(c), st.encode() if st else None, pose.encode() if pose else None,
                           feat.encode() if feat else None)


def capture_enable(on: bool = True):
    lib().lsfm_shim_capture_enable(C.c_int(1 if on else 0))


def capture():
    """Last linear system the reference handed to the CHOLMOD shim (block pattern, ordering,
    scalar CSC, rhs, solution)."""
    cp = lib().lsfm_shim_capture().contents
    out = {}
    if cp.Ap:
        m = cp.m
        Ap = np.ctypeslib.as_array(cp.Ap, shape=(m + 1,)).copy()
        out.update(m=m, Ap=Ap, Ai=np.ctypeslib.as_array(cp.Ai, shape=(max(int(Ap[-1]), 1),))[:Ap[-1]].copy(),
                   bperm=np.ctypeslib.as_array(cp.bperm, shape=(m,)).copy())
    if cp.Sp:
        n = cp.n
        Sp = np.ctypeslib.as_array(cp.Sp, shape=(n + 1,)).copy()
        nz = int(Sp[-1])
        out.update(n=n, Sp=Sp, Si=np.ctypeslib.as_array(cp.Si, shape=(nz,)).copy(),
                   Sx=np.ctypeslib.as_array(cp.Sx, shape=(nz,)).copy())
        if cp.perm:
            out["perm"] = np.ctypeslib.as_array(cp.perm, shape=(n,)).copy()
        if cp.b:
            out["b"] = np.ctypeslib.as_array(cp.b, shape=(n,)).copy()
            out["x"] = np.ctypeslib.as_array(cp.x, shape=(n,)).copy()
    return out


def shim_order(Ap: np.ndarray, Ai: np.ndarray) -> np.ndarray:
    """The oracle's independent implementation of the LSFM-ND block ordering."""
    Ap = _keep(Ap, np.int32); Ai = _keep(Ai, np.int32)
    n = Ap.shape[0] - 1
    perm = np.zeros(n, np.int32)
    lib().lsfm_shim_order(C.c_int(n), Ap.ctypes.data_as(_pi), Ai.ctypes.data_as(_pi),
                          perm.ctypes.data_as(_pi))
    return perm


def shim_times():
    L = lib()
    return dict(amd=L.lsfm_shim_time(0), analyze=L.lsfm_shim_time(1), factorize=L.lsfm_shim_time(2),
                solve=L.lsfm_shim_time(3))


def shim_reset_timers():
    lib().lsfm_shim_reset_timers()


def run_levels_stereo(maps, teacher=None):
    """The reference's merge tree (lmj_PF3D_Divide_ConquerStereo, LinearSFMImp.cpp:1926-2099) driven
    level by level from Python with the reference's OWN operators, yielding every intermediate of a
    level so that a test can feed exactly these inputs to the implementation under test
    ("teacher forcing": errors of one level never reach the next).

    Yields one dict per level:
        level                    level number (0 = leaves)
        E, C                     End / Cur inputs of every pair (lists of LocalMap)
        Et                       End after lmj_Transform_PF3DStereo(End, Cur.Ref)      (1964)
        J                        joined maps, lmj_LinearLS_PF3DStereo(Et, Cur)           (1991)
        rb_idx, rb_in, rb_ref, rb_out   outputs re-based to their first frame           (1997-2025)
        next                     the maps handed to the next level
    and finally {'level': 'final', 'rb_in': [root], 'rb_ref': [FRef], 'rb_out': [result]} (2039-2063).
    """
    level = list(maps)
    L = 0
    while len(level) > 1:
        cnt = len(level)
        npairs = cnt // 2
        E = [level[2 * i] for i in range(npairs)]
        Cm = [level[2 * i + 1] for i in range(npairs)]
        Et = [transform_stereo(e, c.Ref) for e, c in zip(E, Cm)]
        J = [join_stereo(et, c) for et, c in zip(Et, Cm)]
        nxt = list(J)
        if cnt % 2:
            nxt.append(level[cnt - 1])
        rb_idx, rb_in, rb_ref, rb_out = [], [], [], []
        for i in range(len(nxt)):
            if (i + 1) % 2 == 0 and nxt[i].Ref > nxt[i].FRef:
                rb_idx.append(i); rb_in.append(nxt[i]); rb_ref.append(nxt[i].FRef)
                nxt[i] = transform_stereo(nxt[i], nxt[i].FRef)
                rb_out.append(nxt[i])
        yield dict(level=L, E=E, C=Cm, Et=Et, J=J, rb_idx=rb_idx, rb_in=rb_in, rb_ref=rb_ref,
                   rb_out=rb_out, next=nxt)
        level = nxt
        L += 1
    root = level[0]
    if root.Ref > root.FRef:
        out = transform_stereo(root, root.FRef)
        yield dict(level="final", rb_in=[root], rb_ref=[root.FRef], rb_out=[out], next=[out])
    else:
        yield dict(level="final", rb_in=[], rb_ref=[], rb_out=[], next=[root])
