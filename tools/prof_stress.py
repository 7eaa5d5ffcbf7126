"""One warm-up + one solve of a scaled loop-closure scene for ncu launch lists (tools/stress.py without the
reporting).  Usage: python tools/prof_stress.py N landmarks_per_frame [lap]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linearsfm_b200 import api, synth
N = int(sys.argv[1]); fpf = int(sys.argv[2])
lap = int(sys.argv[3]) if len(sys.argv) > 3 else max(500, N // 4)
maps = synth.make_stereo_scene(N, feats_per_frame=fpf, revisit=0.1, lap=lap, max_depth=15.0, gate=True)
api.init(0)
tree = api.Tree(maps)
for _ in range(2):
    tree.reset(); tree.solve()
print("done", tree.result_shape(0).m, tree.last_solve_ms())
