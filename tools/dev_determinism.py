"""Two solves of a loop-closure scene: are the results bit-identical?  Usage: dev_determinism.py N fpf lap"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from linearsfm_b200 import api, synth
N = int(sys.argv[1]); fpf = int(sys.argv[2]); lap = int(sys.argv[3])
maps = synth.make_stereo_scene(N, feats_per_frame=fpf, revisit=0.1, lap=lap, max_depth=15.0, gate=True)
api.init(0)
res = []
for _ in range(2):
    t = api.Tree(maps); t.solve(); res.append(t.download(0)); t.close()
a, b = res
same = all(np.array_equal(getattr(a, f), getattr(b, f)) for f in ("stno", "stVal", "U", "W", "V", "Ui", "Uj", "photo", "feature"))
d = float(np.max(np.abs(a.stVal - b.stVal)) / np.max(np.abs(a.stVal)))
print(f"closed scene N={N} lap={lap}: two solves bit-identical: {same}; state max rel diff {d:.3e}")
