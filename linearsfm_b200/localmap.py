"""Host-side local-map container and the reference's text file formats.

`LocalMap` mirrors the reference's POD containers `LocalMapInfoStereo` / `LocalMapInfo`
(/root/reference/linux/src/LinearSFMImp/LinearSFMImp.h:75-121, 124-178): same field names, same
array shapes, row-major blocks.  The text format is the one parsed by
`lmj_readInformationStereo` (LinearSFMImp.cpp:3044-3132) and `lmj_readInformationMono`
(LinearSFMImp.cpp:6660-6754); see SURVEY.md Appendix A.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


@dataclass
class LocalMap:
    Ref: int
    stno: np.ndarray      # int32 [r]   -poseID x6 rows, then +featID x3 rows
    stVal: np.ndarray     # float64 [r]
    m: int
    n: int
    U: np.ndarray         # float64 [nU, 6, 6]
    Ui: np.ndarray        # int32 [nU]
    Uj: np.ndarray        # int32 [nU]
    W: np.ndarray         # float64 [nW, 6, 3]
    photo: np.ndarray     # int32 [nW]
    feature: np.ndarray   # int32 [nW] non-decreasing
    V: np.ndarray         # float64 [n, 3, 3]
    FBlock: np.ndarray    # int32 [n]
    FRef: int = None
    # mono only (LinearSFMImp.h:172-176)
    ScaP: int = 0
    Fix: int = 0
    Sign: int = 0
    FScaP: int = 0
    FFix: int = 0

    def __post_init__(self):
        if self.FRef is None:
            self.FRef = self.Ref
        self.stno = np.ascontiguousarray(self.stno, dtype=np.int32).reshape(-1)
        self.stVal = np.ascontiguousarray(self.stVal, dtype=np.float64).reshape(-1)
        self.U = np.ascontiguousarray(self.U, dtype=np.float64).reshape(-1, 6, 6)
        self.Ui = np.ascontiguousarray(self.Ui, dtype=np.int32).reshape(-1)
        self.Uj = np.ascontiguousarray(self.Uj, dtype=np.int32).reshape(-1)
        self.W = np.ascontiguousarray(self.W, dtype=np.float64).reshape(-1, 6, 3)
        self.photo = np.ascontiguousarray(self.photo, dtype=np.int32).reshape(-1)
        self.feature = np.ascontiguousarray(self.feature, dtype=np.int32).reshape(-1)
        self.V = np.ascontiguousarray(self.V, dtype=np.float64).reshape(-1, 3, 3)
        self.FBlock = np.ascontiguousarray(self.FBlock, dtype=np.int32).reshape(-1)

    @property
    def r(self) -> int:
        return 6 * self.m + 3 * self.n

    @property
    def nU(self) -> int:
        return int(self.Ui.shape[0])

    @property
    def nW(self) -> int:
        return int(self.photo.shape[0])

    def pose_ids(self) -> np.ndarray:
        return -self.stno[0:6 * self.m:6]

    def feature_ids(self) -> np.ndarray:
        return self.stno[6 * self.m::3]

    def poses(self) -> np.ndarray:
        return self.stVal[:6 * self.m].reshape(self.m, 6)

    def features(self) -> np.ndarray:
        return self.stVal[6 * self.m:].reshape(self.n, 3)

    def dense_information(self) -> np.ndarray:
        """Full symmetric information matrix (small maps only; test helper)."""
        r = self.r
        I = np.zeros((r, r))
        for b in range(self.nU):
            i, j = int(self.Ui[b]), int(self.Uj[b])
            blk = self.U[b]
            if i == j:
                I[6 * i:6 * i + 6, 6 * i:6 * i + 6] += blk
            else:
                I[6 * i:6 * i + 6, 6 * j:6 * j + 6] += blk
                I[6 * j:6 * j + 6, 6 * i:6 * i + 6] += blk.T
        o = 6 * self.m
        for b in range(self.nW):
            p, f = int(self.photo[b]), int(self.feature[b])
            I[6 * p:6 * p + 6, o + 3 * f:o + 3 * f + 3] += self.W[b]
            I[o + 3 * f:o + 3 * f + 3, 6 * p:6 * p + 6] += self.W[b].T
        for f in range(self.n):
            I[o + 3 * f:o + 3 * f + 3, o + 3 * f:o + 3 * f + 3] += self.V[f]
        return I


def _fmt(a: np.ndarray) -> str:
    return " ".join(repr(float(x)) for x in np.asarray(a).reshape(-1))


def write_localmap(path: str, lm: LocalMap, mono: bool = False) -> None:
    """Write one `localmap_<i>.txt`. Doubles are written with
    repr() so they round-trip bit-exactly through the reference's fscanf("%lf")."""
    with open(path, "w") as f:
        if mono:
            f.write(f"{lm.Ref} {lm.ScaP} {lm.Fix} {lm.Sign}\n{lm.r}\n")
        else:
            f.write(f"{lm.Ref}\n{lm.r}\n")
        f.write("\n".join(f"{int(s)} {float(v)!r}" for s, v in zip(lm.stno, lm.stVal)))
        f.write(f"\n{lm.m} {lm.n}\n{lm.nU}\n")
        f.write(_fmt(lm.U) + "\n")
        f.write(" ".join(str(int(x)) for x in lm.Ui) + "\n")
        f.write(" ".join(str(int(x)) for x in lm.Uj) + "\n")
        f.write(f"{lm.nW}\n")
        f.write(_fmt(lm.W) + "\n")
        f.write(" ".join(str(int(x)) for x in lm.photo) + "\n")
        f.write(" ".join(str(int(x)) for x in lm.feature) + "\n")
        f.write(_fmt(lm.V) + "\n")
        f.write(" ".join(str(int(x)) for x in lm.FBlock) + "\n")


def read_localmap(path: str, mono: bool = False) -> LocalMap:
    """Python twin of the reference readers (test helper; the product parser is C++)."""
    with open(path) as f:
        tok = f.read().split()
    it = iter(tok)
    Ref = int(next(it))
    ScaP = Fix = Sign = 0
    if mono:
        ScaP, Fix, Sign = int(next(it)), int(next(it)), int(next(it))
    r = int(next(it))
    stno = np.empty(r, np.int32)
    stVal = np.empty(r, np.float64)
    for i in range(r):
        stno[i] = int(next(it))
        stVal[i] = float(next(it))
    m, n = int(next(it)), int(next(it))
    nU = int(next(it))
    U = np.array([float(next(it)) for _ in range(36 * nU)])
    Ui = np.array([int(next(it)) for _ in range(nU)], np.int32)
    Uj = np.array([int(next(it)) for _ in range(nU)], np.int32)
    nW = int(next(it))
    W = np.array([float(next(it)) for _ in range(18 * nW)])
    photo = np.array([int(next(it)) for _ in range(nW)], np.int32)
    feature = np.array([int(next(it)) for _ in range(nW)], np.int32)
    V = np.array([float(next(it)) for _ in range(9 * n)])
    FBlock = np.array([int(next(it)) for _ in range(n)], np.int32)
    lm = LocalMap(Ref=Ref, stno=stno, stVal=stVal, m=m, n=n, U=U, Ui=Ui, Uj=Uj, W=W, photo=photo,
                  feature=feature, V=V, FBlock=FBlock, ScaP=ScaP, Fix=Fix, Sign=Sign)
    if mono:
        lm.FScaP, lm.FFix = ScaP, Fix
    return lm


def maps_equal_int(a: LocalMap, b: LocalMap) -> list[str]:
    """Names of integer fields that differ (bit-exact comparison)."""
    bad = []
    for k in ("Ref", "FRef", "m", "n"):
        if int(getattr(a, k)) != int(getattr(b, k)):
            bad.append(k)
    for k in ("stno", "Ui", "Uj", "photo", "feature", "FBlock"):
        x, y = getattr(a, k), getattr(b, k)
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append(k)
    return bad
