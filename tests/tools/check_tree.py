"""Run one stereo merge tree on the GPU and compare with the oracle (used by tests that need a fresh
process, e.g. with LSFM_FORCE_OVERFLOW=1 / LSFM_TF_V3=1 which are read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from linearsfm_b200 import api, synth  # noqa: E402
import ref_oracle as ro  # noqa: E402
from util import assert_maps_match, state_rel_err  # noqa: E402

n, fpf = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3].startswith("closed"):      # loop closures every <k> frames: closed<k>
    maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=100 + n, revisit=0.1, lap=int(sys.argv[3][6:]),
                                   max_depth=15.0, gate=True)
else:
    maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=100 + n)
ref, _, _ = ro.run_tree_stereo(maps)
api.init(0)
got = api.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(maps)
assert_maps_match(got, ref, tol_state=1e-7, tol_info=1e-7, what=f"tree N={n}")
e = state_rel_err(got, ref)
assert e <= 1e-6
print(f"check_tree ok N={n} state rel err {e:.2e}")
