"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).  Run in the build
container (needs /root/reference to build the oracle):  python tests/golden/make_golden.py
Each fixture stores the inputs and the reference's outputs of one operator / one whole tree so the
GPU path can be checked on a box where the reference sources do not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_oracle as ro  # noqa: E402
from linearsfm_b200 import synth  # noqa: E402

FIELDS = ("stno", "stVal", "U", "Ui", "Uj", "W", "photo", "feature", "V", "FBlock")


def pack(prefix, lm, d):
    for f in FIELDS:
        d[f"{prefix}_{f}"] = getattr(lm, f)
    d[f"{prefix}_meta"] = np.array([lm.Ref, lm.FRef, lm.m, lm.n], np.int64)


def main():
    ro.build()
    maps = synth.make_stereo_scene(5, feats_per_frame=6, seed=424242)
    d = {}
    for i, lm in enumerate(maps):
        pack(f"leaf{i}", lm, d)
    t0 = ro.transform_stereo(maps[0], maps[1].Ref)
    pack("tf0", t0, d)
    j0 = ro.join_stereo(t0, maps[1])
    pack("join0", j0, d)
    t2 = ro.transform_stereo(maps[2], maps[3].Ref)
    j1 = ro.join_stereo(t2, maps[3])
    j1b = ro.transform_stereo(j1, j1.FRef)
    pack("rebase1", j1b, d)
    fin, _, _ = ro.run_tree_stereo(maps)
    pack("final", fin, d)
    np.savez_compressed(os.path.join(HERE, "stereo_n5.npz"), **d)
    print("wrote stereo_n5.npz", {k: v.shape for k, v in list(d.items())[:3]}, "final m,n", fin.m, fin.n)


if __name__ == "__main__":
    main()


def main_mono():
    ro.build()
    maps = synth.make_mono_scene(5, feats_per_frame=8, seed=515151)
    d = {}
    meta = lambda lm: np.array([lm.Ref, lm.FRef, lm.m, lm.n, lm.ScaP, lm.Fix, lm.Sign, lm.FScaP, lm.FFix], np.int64)
    def pk(prefix, lm):
        for f in FIELDS:
            d[f"{prefix}_{f}"] = getattr(lm, f)
        d[f"{prefix}_meta"] = meta(lm)
    for i, lm in enumerate(maps):
        pk(f"leaf{i}", lm)
    t0 = ro.transform_mono(maps[0], maps[1].Ref, maps[1].ScaP, maps[1].Fix)
    pk("tf0", t0)
    pk("join0", ro.join_mono(t0, maps[1]))
    fin, _, _ = ro.run_tree_mono(maps)
    pk("final", fin)
    np.savez_compressed(os.path.join(HERE, "mono_n5.npz"), **d)
    print("wrote mono_n5.npz final m,n", fin.m, fin.n)


if __name__ == "__main__":
    main_mono()


def main_root_pattern():
    """Block pattern (upper CSC, as the reference hands it to cholmod_amd, LinearSFMImp.cpp:2529-2549)
    of the reduced camera system of the ROOT join of the 3499-map headline scene, captured from the
    reference through the shim (~1 minute of CPU)."""
    ro.build()
    maps = synth.make_stereo_scene(3499, feats_per_frame=128)
    ro.capture_enable(True)
    ro.run_tree_stereo(maps)
    c = ro.capture()
    ro.capture_enable(False)
    np.savez_compressed(os.path.join(HERE, "root_pattern_3499.npz"), Ap=c["Ap"].astype(np.int32),
                        Ai=c["Ai"].astype(np.int32))
    print("wrote root_pattern_3499.npz m", c["m"], "blocks", len(c["Ai"]))


if __name__ == "__main__" and "--root-pattern" in __import__("sys").argv:
    main_root_pattern()


def main_root_pattern_closed():
    """Block pattern of the reduced camera system of the ROOT join of the 3499-map scene with loop
    closures (synth revisit=0.1, lap=500, max_depth=15, gate): pose pairs that share a feature of the
    reference's root joint map + its U pattern (the reference's smask rule, LinearSFMImp.cpp:2156-2173),
    upper CSC.  The reference's tree is run level by level up to the root (several minutes of CPU)."""
    import scipy.sparse as sp
    ro.build()
    maps = synth.make_stereo_scene(3499, feats_per_frame=128, revisit=0.1, lap=500, max_depth=15.0, gate=True)
    j = None
    for rec in ro.run_levels_stereo(maps):
        if rec["level"] != "final":
            j = rec["J"][0]
    m = j.m
    A = sp.csr_matrix((np.ones(len(j.feature)), (j.photo, j.feature)), shape=(m, j.n))
    Pm = (A @ A.T).tocoo()
    rows = np.concatenate([Pm.row, j.Ui, j.Uj]); cols = np.concatenate([Pm.col, j.Uj, j.Ui])
    G = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(m, m))
    U = sp.triu(G).tocsc(); U.sort_indices()
    np.savez_compressed(os.path.join(HERE, "root_pattern_closed_3499.npz"), Ap=U.indptr.astype(np.int32),
                        Ai=U.indices.astype(np.int32))
    print("wrote root_pattern_closed_3499.npz", m, U.nnz)
