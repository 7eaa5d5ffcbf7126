"""One merge-tree solve (after one warm-up solve) for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linearsfm_b200 import api, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3499
if len(sys.argv) > 3 and sys.argv[3] == "closed":      # loop closures every 500 frames (bench.py --scene closed)
    maps = synth.make_stereo_scene(N, feats_per_frame=128, revisit=0.1, lap=500, max_depth=15.0, gate=True)
else:
    maps = synth.make_stereo_scene(N, feats_per_frame=128)
api.init(0)
tree = api.Tree(maps)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    tree.solve()
print("done", tree.result_shape(0).m)
