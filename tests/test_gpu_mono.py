"""Monocular path: CUDA vs the reference's mono operators on the synthetic
RS-shape scene (3 frames per map, two shared poses)."""
import numpy as np
import pytest

from linearsfm_b200 import synth
from util import assert_maps_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mono8():
    return synth.make_mono_scene(8, feats_per_frame=20, seed=77)


def check_meta(got, ref):
    for k in ("ScaP", "Fix", "Sign", "FScaP", "FFix"):
        assert getattr(got, k) == getattr(ref, k), k


def test_mono_transform_leaf(gpu, oracle, mono8):
    for k in (0, 3):
        nxt = mono8[k + 1]
        ref = oracle.transform_mono(mono8[k], nxt.Ref, nxt.ScaP, nxt.Fix)
        got = gpu.transform_mono_batch([mono8[k]], [nxt.Ref], [nxt.ScaP], [nxt.Fix])[0]
        assert_maps_match(got, ref, what=f"mono leaf transform {k}")
        check_meta(got, ref)


def test_mono_transform_joint(gpu, oracle, mono8):
    e = oracle.transform_mono(mono8[2], mono8[3].Ref, mono8[3].ScaP, mono8[3].Fix)
    j = oracle.join_mono(e, mono8[3])
    assert j.Ref > j.FRef
    ref = oracle.transform_mono(j, j.FRef, j.FScaP, j.FFix)
    got = gpu.transform_mono_batch([j], [j.FRef], [j.FScaP], [j.FFix])[0]
    assert_maps_match(got, ref, what="mono re-base of a joint map")
    check_meta(got, ref)
    # and a level-2 style transform: joint of maps 1-2 into the frame of the re-based joint of 3-4
    e0 = oracle.transform_mono(mono8[0], mono8[1].Ref, mono8[1].ScaP, mono8[1].Fix)
    j0 = oracle.join_mono(e0, mono8[1])
    ref2 = oracle.transform_mono(j0, ref.Ref, ref.ScaP, ref.Fix)
    got2 = gpu.transform_mono_batch([j0], [ref.Ref], [ref.ScaP], [ref.Fix])[0]
    assert_maps_match(got2, ref2, what="mono transform of a joint map")
