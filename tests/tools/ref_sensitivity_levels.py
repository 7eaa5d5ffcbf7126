"""Where does the reference's own sensitivity to 1e-15 input noise come from?  Runs the oracle's merge tree
level by level on a scene and on a perturbed copy and prints, per level, how far the two runs are apart.
Usage: python tests/tools/ref_sensitivity_levels.py N [landmarks_per_frame [revisit [lap [max_depth [gate]]]]]"""
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
md = float(sys.argv[5]) if len(sys.argv) > 5 else 30.0
gate = bool(int(sys.argv[6])) if len(sys.argv) > 6 else False
maps = synth.make_stereo_scene(N, feats_per_frame=fpf, revisit=rev, lap=lap, max_depth=md, gate=gate)
rng = np.random.default_rng(0)
pert = []
for m in maps:
    m2 = copy.deepcopy(m); m2.W = m2.W * (1 + 1e-15 * rng.standard_normal(m2.W.shape)); pert.append(m2)
for a, b in zip(ro.run_levels_stereo(maps), ro.run_levels_stereo(pert)):
    if a["level"] == "final":
        print("final", rel_err(a["next"][0].stVal, b["next"][0].stVal))
        break
    ej = max(rel_err(x.stVal, y.stVal) for x, y in zip(a["J"], b["J"]))
    et = max(rel_err(x.W, y.W) for x, y in zip(a["Et"], b["Et"]))
    en = max(rel_err(x.stVal, y.stVal) for x, y in zip(a["next"], b["next"]))
    ew = max(rel_err(x.W, y.W) for x, y in zip(a["next"], b["next"]))
    print("level", a["level"], "pairs", len(a["J"]), "m", a["J"][0].m, "Et.W %.2e  J.state %.2e  next.state %.2e next.W %.2e" % (et, ej, en, ew), flush=True)
