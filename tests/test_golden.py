"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py).
CPU: the oracle built here still reproduces them (guards the oracle build and the scene
generator).  GPU: the CUDA path matches them without needing the reference sources."""
import os

import numpy as np
import pytest

from linearsfm_b200.localmap import LocalMap
from util import assert_maps_match

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stereo_n5.npz")


def load(prefix, d):
    meta = d[f"{prefix}_meta"]
    return LocalMap(Ref=int(meta[0]), FRef=int(meta[1]), m=int(meta[2]), n=int(meta[3]),
                    **{f: d[f"{prefix}_{f}"] for f in ("stno", "stVal", "U", "Ui", "Uj", "W", "photo",
                                                       "feature", "V", "FBlock")})


@pytest.fixture(scope="module")
def gold():
    d = np.load(G)
    return d, [load(f"leaf{i}", d) for i in range(5)]


def test_generator_is_stable(gold):
    from linearsfm_b200 import synth
    d, leaves = gold
    maps = synth.make_stereo_scene(5, feats_per_frame=6, seed=424242)
    for a, b in zip(maps, leaves):
        assert_maps_match(a, b, tol_state=1e-13, tol_info=1e-12, what="synthetic leaf")


def test_oracle_reproduces_golden(oracle, gold):
    d, leaves = gold
    t0 = oracle.transform_stereo(leaves[0], leaves[1].Ref)
    assert_maps_match(t0, load("tf0", d), tol_state=1e-14, tol_info=1e-13, what="tf0")
    assert_maps_match(oracle.join_stereo(t0, leaves[1]), load("join0", d), tol_state=1e-12, tol_info=1e-13, what="join0")
    fin, _, _ = oracle.run_tree_stereo(leaves)
    assert_maps_match(fin, load("final", d), tol_state=1e-11, tol_info=1e-11, what="final")


@pytest.mark.gpu
def test_gpu_matches_golden(gpu, gold):
    d, leaves = gold
    t0 = gpu.transform_stereo_batch([leaves[0]], [leaves[1].Ref])[0]
    assert_maps_match(t0, load("tf0", d), what="tf0")
    j0 = gpu.join_stereo_batch([load("tf0", d)], [leaves[1]])[0]
    assert_maps_match(j0, load("join0", d), what="join0")
    fin = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(leaves)
    assert_maps_match(fin, load("final", d), tol_state=1e-8, tol_info=1e-8, what="final")


GM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mono_n5.npz")


def load_mono(prefix, d):
    mt = d[f"{prefix}_meta"]
    return LocalMap(Ref=int(mt[0]), FRef=int(mt[1]), m=int(mt[2]), n=int(mt[3]), ScaP=int(mt[4]), Fix=int(mt[5]),
                    Sign=int(mt[6]), FScaP=int(mt[7]), FFix=int(mt[8]),
                    **{f: d[f"{prefix}_{f}"] for f in ("stno", "stVal", "U", "Ui", "Uj", "W", "photo",
                                                       "feature", "V", "FBlock")})


def test_oracle_reproduces_mono_golden(oracle):
    d = np.load(GM)
    leaves = [load_mono(f"leaf{i}", d) for i in range(5)]
    t0 = oracle.transform_mono(leaves[0], leaves[1].Ref, leaves[1].ScaP, leaves[1].Fix)
    assert_maps_match(t0, load_mono("tf0", d), tol_state=1e-13, tol_info=1e-12, what="mono tf0")
    fin, _, _ = oracle.run_tree_mono(leaves)
    assert_maps_match(fin, load_mono("final", d), tol_state=1e-10, tol_info=1e-10, what="mono final")


@pytest.mark.gpu
def test_gpu_matches_mono_golden(gpu):
    d = np.load(GM)
    leaves = [load_mono(f"leaf{i}", d) for i in range(5)]
    t0 = gpu.transform_mono_batch([leaves[0]], [leaves[1].Ref], [leaves[1].ScaP], [leaves[1].Fix])[0]
    assert_maps_match(t0, load_mono("tf0", d), what="mono tf0")
    j0 = gpu.join_mono_batch([load_mono("tf0", d)], [leaves[1]])[0]
    assert_maps_match(j0, load_mono("join0", d), what="mono join0")
    fin = gpu.run_mono(leaves)
    assert_maps_match(fin, load_mono("final", d), tol_state=1e-8, tol_info=1e-8, what="mono final")
