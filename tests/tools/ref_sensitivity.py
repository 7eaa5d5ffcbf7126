"""Conditioning of the REFERENCE on a synthetic stereo scene: run the oracle (unmodified reference TU)
on the scene and on a copy whose W blocks carry 1e-15 relative noise; print how far its own final
state / information blocks move.  Measured here (seed of synth.make_stereo_scene, 128 landmarks/frame):
    N= 466: state 4.6e-09  N=1200: state 2.8e-06, U/W 7.9e-06     N=3499: state 1.6e-03, U/W 2.2e-03
With loop closures (revisit=0.1, lap=500) the chain is anchored and the same measure stays small at 3499 maps.
Usage: python tests/tools/ref_sensitivity.py N [landmarks_per_frame [revisit [lap]]]"""
import copy, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from linearsfm_b200 import synth  # noqa: E402
import ref_oracle as ro  # noqa: E402
from util import rel_err  # noqa: E402

N = int(sys.argv[1]); fpf = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rev = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
lap = int(sys.argv[4]) if len(sys.argv) > 4 else 500
maps = synth.make_stereo_scene(N, feats_per_frame=fpf, revisit=rev, lap=lap)
ref, _, _ = ro.run_tree_stereo(maps)
rng = np.random.default_rng(0)
pert = []
for m in maps:
    m2 = copy.deepcopy(m); m2.W = m2.W * (1 + 1e-15 * rng.standard_normal(m2.W.shape)); pert.append(m2)
ref2, _, _ = ro.run_tree_stereo(pert)
print(N, fpf, rev, lap, "m n nU nW", ref.m, ref.n, ref.nU, ref.nW, "reference self-sensitivity: stVal %.3e U %.3e W %.3e V %.3e" % (
    rel_err(ref.stVal, ref2.stVal), rel_err(ref.U, ref2.U), rel_err(ref.W, ref2.W), rel_err(ref.V, ref2.V)))
