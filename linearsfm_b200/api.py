"""Python host-side mirror of the reference's operator interface for the hot path.

The reference's in-process boundary is the public methods of `class CLinearSFMImp`
(/root/reference/linux/src/LinearSFMImp/LinearSFMImp.h:181-253).  `CLinearSFMImp` below keeps the
same method names and argument meaning for the stereo path and forwards every call to the C ABI
(include/linearsfm_b200.h); nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import LsfmMap, LsfmError, lib, check
from .localmap import LocalMap

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def to_c(lm: LocalMap):
    ka = dict(stno=_c(lm.stno, np.int32), stVal=_c(lm.stVal, np.float64), U=_c(lm.U, np.float64),
              Ui=_c(lm.Ui, np.int32), Uj=_c(lm.Uj, np.int32), W=_c(lm.W, np.float64),
              photo=_c(lm.photo, np.int32), feature=_c(lm.feature, np.int32),
              V=_c(lm.V, np.float64), FBlock=_c(lm.FBlock, np.int32))
    c = LsfmMap(Ref=lm.Ref, FRef=lm.FRef, r=lm.r, m=lm.m, n=lm.n, nU=lm.nU, nW=lm.nW,
                ScaP=lm.ScaP, Fix=lm.Fix, Sign=lm.Sign, FScaP=lm.FScaP, FFix=lm.FFix)
    for k, v in ka.items():
        setattr(c, k, v.ctypes.data_as(_pi if v.dtype == np.int32 else _pd))
    return c, ka


def to_c_array(maps):
    arr = (LsfmMap * len(maps))()
    keep = []
    for i, lm in enumerate(maps):
        c, ka = to_c(lm)
        arr[i] = c
        keep.append(ka)
    return arr, keep


def from_c(c: LsfmMap, free: bool = True) -> LocalMap:
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
        lib().lsfm_free_map(C.byref(c))
    return lm


def init(device: int = 0):
    check(lib().lsfm_init(C.c_int(device)))


def device_count() -> int:
    return int(lib().lsfm_device_count())


def stats_reset(stage_timing: bool = False, objective: bool = False):
    lib().lsfm_stats_reset(C.c_int((1 if stage_timing else 0) | (2 if objective else 0)))


def stats() -> dict:
    return json.loads(lib().lsfm_stats_json().decode())


def transform_stereo_batch(maps, refs):
    arr, _keep = to_c_array(maps)
    out = (LsfmMap * len(maps))()
    r = (C.c_int * len(maps))(*[int(x) for x in refs])
    check(lib().lsfm_transform_stereo_batch(arr, r, C.c_int(len(maps)), out))
    return [from_c(out[i]) for i in range(len(maps))]


def transform_mono_batch(maps, refs, scaps, fixes):
    arr, _keep = to_c_array(maps)
    out = (LsfmMap * len(maps))()
    n = len(maps)
    r = (C.c_int * n)(*[int(x) for x in refs])
    sc = (C.c_int * n)(*[int(x) for x in scaps])
    fx = (C.c_int * n)(*[int(x) for x in fixes])
    check(lib().lsfm_transform_mono_batch(arr, r, sc, fx, C.c_int(n), out))
    return [from_c(out[i]) for i in range(n)]


def join_stereo_batch(ends, curs):
    a, _k1 = to_c_array(ends)
    b, _k2 = to_c_array(curs)
    out = (LsfmMap * len(ends))()
    check(lib().lsfm_join_stereo_batch(a, b, C.c_int(len(ends)), out))
    return [from_c(out[i]) for i in range(len(ends))]


def join_mono_batch(ends, curs):
    a, _k1 = to_c_array(ends)
    b, _k2 = to_c_array(curs)
    out = (LsfmMap * len(ends))()
    check(lib().lsfm_join_mono_batch(a, b, C.c_int(len(ends)), out))
    return [from_c(out[i]) for i in range(len(ends))]


def run_mono(maps) -> LocalMap:
    """lmj_PF3D_Divide_ConquerMono (LinearSFMImp.cpp:6511): whole mono merge tree, host maps in/out."""
    arr, _keep = to_c_array(maps)
    out = LsfmMap()
    check(lib().lsfm_run_mono(arr, C.c_int(len(maps)), C.byref(out)))
    return from_c(out)


def block_ordering(Ap, Ai) -> np.ndarray:
    Ap = _c(Ap, np.int32); Ai = _c(Ai, np.int32)
    m = Ap.shape[0] - 1
    perm = np.zeros(m, np.int32)
    check(lib().lsfm_block_ordering(C.c_int(m), Ap.ctypes.data_as(_pi), Ai.ctypes.data_as(_pi),
                                    perm.ctypes.data_as(_pi)))
    return perm


def debug_last_solve() -> dict:
    m = C.c_int(0)
    rp, ci, pm = _pi(), _pi(), _pi()
    S, E = _pd(), _pd()
    check(lib().lsfm_debug_last_solve(C.byref(m), C.byref(rp), C.byref(ci), C.byref(S), C.byref(E),
                                      C.byref(pm)))
    mm = m.value
    rowptr = np.ctypeslib.as_array(rp, shape=(mm + 1,)).copy()
    nz = int(rowptr[-1])
    return dict(m=mm, rowptr=rowptr, colidx=np.ctypeslib.as_array(ci, shape=(nz,)).copy(),
                S=np.ctypeslib.as_array(S, shape=(nz, 6, 6)).copy(),
                E=np.ctypeslib.as_array(E, shape=(6 * mm,)).copy(),
                perm=np.ctypeslib.as_array(pm, shape=(mm,)).copy())


def pcg_block(rowptr, colidx, S, E, tol=1e-13, max_iters=None):
    """Block-Jacobi PCG on the reduced camera system (cross-check of the Cholesky): returns
    (x, iterations, relative residual)."""
    rowptr = _c(rowptr, np.int32); colidx = _c(colidx, np.int32)
    S = _c(S, np.float64); E = _c(E, np.float64)
    m = rowptr.shape[0] - 1
    x = np.zeros(6 * m)
    it = C.c_int(0); rr = C.c_double(0.0)
    check(lib().lsfm_pcg_block(C.c_int(m), rowptr.ctypes.data_as(_pi), colidx.ctypes.data_as(_pi),
                               S.ctypes.data_as(_pd), E.ctypes.data_as(_pd), x.ctypes.data_as(_pd),
                               C.c_double(tol), C.c_int(max_iters or 40 * 6 * m), C.byref(it), C.byref(rr)))
    return x, it.value, rr.value


def _cov_blocks(m, sel, flat):
    out, o = [], 0
    for b in sel:
        d = 6 if b < m else 3
        out.append(flat[o:o + d * d].reshape(d, d).copy())
        o += d * d
    return out


def marginal_cov(lm: LocalMap, sel):
    """Marginal covariance blocks (6x6 per pose index in [0, m), 3x3 per feature index m + f) of the inverse of the
    map's information matrix, through the joint solver with unit right-hand sides (lsfm_marginal_cov_stereo)."""
    sel = _c(sel, np.int32)
    flat = np.zeros(int(sum(36 if b < lm.m else 9 for b in sel)))
    c, _keep = to_c(lm)
    check(lib().lsfm_marginal_cov_stereo(C.byref(c), C.c_int(len(sel)), sel.ctypes.data_as(_pi),
                                         flat.ctypes.data_as(_pd)))
    return _cov_blocks(lm.m, sel, flat)


class Tree:
    """Leaf maps resident in HBM; `solve()` runs the merge tree on the device."""

    def __init__(self, maps):
        arr, keep = to_c_array(maps)
        self._h = C.c_void_p()
        check(lib().lsfm_tree_create_stereo(arr, C.c_int(len(maps)), C.byref(self._h)))
        self.num = len(maps)

    def set_maps(self, maps):
        arr, keep = to_c_array(maps)
        self.set_maps_c(arr, len(maps))

    def set_maps_c(self, arr, num):
        """arr: prebuilt (LsfmMap * num) array whose pointers stay alive in the caller."""
        check(lib().lsfm_tree_set_maps(self._h, arr, C.c_int(num)))
        self.num = num

    def append_maps(self, maps):
        arr, keep = to_c_array(maps)
        check(lib().lsfm_tree_append_maps(self._h, arr, C.c_int(len(maps))))
        self.num += len(maps)

    def export_device(self, idx: int, dst_ptr: int, nbytes: int):
        """pack result map `idx` into the contiguous device buffer at dst_ptr (D2D copies)."""
        check(lib().lsfm_tree_export_device(self._h, C.c_int(idx), C.c_void_p(dst_ptr), C.c_size_t(nbytes)))

    def append_device(self, shape: LsfmMap, src_ptr: int, nbytes: int):
        """append a map packed by export_device (possibly on another GPU) to the input set."""
        check(lib().lsfm_tree_append_device(self._h, C.byref(shape), C.c_void_p(src_ptr), C.c_size_t(nbytes)))
        self.num += 1

    def reset(self):
        """input set := the maps of the last upload (still resident in HBM)."""
        check(lib().lsfm_tree_reset(self._h))

    def adopt_result(self):
        check(lib().lsfm_tree_adopt_result(self._h))

    def last_solve_ms(self) -> float:
        return float(lib().lsfm_tree_last_solve_ms(self._h))

    def solve(self, verbose: bool = False, first_index: int = 0, max_levels: int = -1):
        check(lib().lsfm_tree_solve(self._h, C.c_int(1 if verbose else 0), C.c_int(first_index),
                                    C.c_int(max_levels)))

    def result_count(self) -> int:
        return int(lib().lsfm_tree_result_count(self._h))

    def result_shape(self, idx: int = 0) -> LsfmMap:
        s = LsfmMap()
        check(lib().lsfm_tree_result_shape(self._h, C.c_int(idx), C.byref(s)))
        return s

    def download(self, idx: int = 0) -> LocalMap:
        out = LsfmMap()
        check(lib().lsfm_tree_download(self._h, C.c_int(idx), C.byref(out)))
        return from_c(out)

    def download_state(self, idx: int = 0, stno=None, stVal=None):
        s = self.result_shape(idx)
        if stno is None:
            stno = np.empty(s.r, np.int32)
        if stVal is None:
            stVal = np.empty(s.r, np.float64)
        check(lib().lsfm_tree_download_state(self._h, C.c_int(idx), stno.ctypes.data_as(_pi),
                                             stVal.ctypes.data_as(_pd)))
        return stno, stVal

    def marginal_cov(self, sel, idx: int = 0):
        """Marginal covariance blocks of result `idx` (resident in HBM), see `marginal_cov`."""
        m = self.result_shape(idx).m
        sel = _c(sel, np.int32)
        flat = np.zeros(int(sum(36 if b < m else 9 for b in sel)))
        check(lib().lsfm_tree_marginal_cov(self._h, C.c_int(idx), C.c_int(len(sel)), sel.ctypes.data_as(_pi),
                                           flat.ctypes.data_as(_pd)))
        return _cov_blocks(m, sel, flat)

    def close(self):
        if self._h:
            lib().lsfm_tree_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CLinearSFMImp:
    """Same public method names as the reference class (LinearSFMImp.h:181-253), stereo path.

    Differences forced by Python: results are returned instead of written through reference
    arguments / the m_GMapS member, and inputs are not freed.
    """

    def __init__(self, device: int = 0):
        init(device)                                   # CLinearSFMImp::CLinearSFMImp (Imp.cpp:80-87)
        self.m_GMapS: LocalMap | None = None

    def lmj_Transform_PF3DStereo(self, GMap: LocalMap, Ref: int) -> LocalMap:
        """LinearSFMImp.cpp:349: re-express `GMap` (the reference's m_GMapS) in the frame of pose Ref."""
        return transform_stereo_batch([GMap], [Ref])[0]

    def lmj_LinearLS_PF3DStereo(self, GMap_End: LocalMap, GMap_Cur: LocalMap) -> LocalMap:
        """LinearSFMImp.cpp:2551: join End (already in Cur's frame) with Cur; result also in m_GMapS."""
        self.m_GMapS = join_stereo_batch([GMap_End], [GMap_Cur])[0]
        return self.m_GMapS

    def lmj_solveLinearSFMStereo(self, eb, ea, U, W, V, Ui, Uj, photo, feature, m, n, nU=None,
                                 nW=None) -> np.ndarray:
        """LinearSFMImp.cpp:2119, same argument order after the output stVal (returned)."""
        ea = _c(ea, np.float64); eb = _c(eb, np.float64)
        U = _c(U, np.float64); W = _c(W, np.float64); V = _c(V, np.float64)
        Ui = _c(Ui, np.int32); Uj = _c(Uj, np.int32)
        photo = _c(photo, np.int32); feature = _c(feature, np.int32)
        nU = Ui.shape[0] if nU is None else nU
        nW = photo.shape[0] if nW is None else nW
        st = np.zeros(6 * m + 3 * n)
        p = lambda a: a.ctypes.data_as(_pi if a.dtype == np.int32 else _pd)
        check(lib().lsfm_solve_stereo(p(st), p(eb), p(ea), p(U), p(W), p(V), p(Ui), p(Uj), p(photo),
                                      p(feature), C.c_int(m), C.c_int(n), C.c_int(nU), C.c_int(nW)))
        return st

    def lmj_solveLinearSFMMono(self, eb, ea, U, W, V, Ui, Uj, photo, feature, m, n, Ref, ScaP, Fix, Sign,
                               FixBlk=0) -> np.ndarray:
        """LinearSFMImp.cpp:6756, same argument order after the output stVal (returned); nU / nW come
        from the array lengths."""
        ea = _c(ea, np.float64); eb = _c(eb, np.float64)
        U = _c(U, np.float64); W = _c(W, np.float64); V = _c(V, np.float64)
        Ui = _c(Ui, np.int32); Uj = _c(Uj, np.int32)
        photo = _c(photo, np.int32); feature = _c(feature, np.int32)
        st = np.zeros(6 * m + 3 * n)
        p = lambda a: a.ctypes.data_as(_pi if a.dtype == np.int32 else _pd)
        check(lib().lsfm_solve_mono(p(st), p(eb), p(ea), p(U), p(W), p(V), p(Ui), p(Uj), p(photo), p(feature),
                                    C.c_int(m), C.c_int(n), C.c_int(Ui.shape[0]), C.c_int(photo.shape[0]),
                                    C.c_int(Ref), C.c_int(ScaP), C.c_int(Fix), C.c_int(Sign), C.c_int(FixBlk)))
        return st

    def lmj_PF3D_Divide_ConquerStereo(self, maps) -> LocalMap:
        """LinearSFMImp.cpp:1926: the whole merge tree, host maps in, final map out."""
        arr, _keep = to_c_array(maps)
        out = LsfmMap()
        check(lib().lsfm_run_stereo(arr, C.c_int(len(maps)), C.byref(out)))
        self.m_GMapS = from_c(out)
        return self.m_GMapS
