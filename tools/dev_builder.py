"""Developer script: raw stereo observations -> CUDA local-map builder -> merge tree, with timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from linearsfm_b200 import api, builder
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3499
fpf = int(sys.argv[2]) if len(sys.argv) > 2 else 128
t = time.time(); pairs, cam, truth = builder.make_stereo_observations(N, feats_per_frame=fpf); print("gen %.1fs" % (time.time() - t))
api.init(0)
for rep in range(3):
    api.stats_reset(stage_timing=True)
    t = time.time(); maps, iters = builder.build_localmaps_stereo(pairs, cam); dt = time.time() - t
    st = api.stats()["stages"].get("builder", {})
    nl = sum(m.n for m in maps)
    print("build %d maps / %d landmarks: wall %.3fs (incl. host packing and D2H), kernel %.3f ms, iterations max %d median %d"
          % (N, nl, dt, st.get("ms", float("nan")), iters.max(), int(np.median(iters))))
terr = max(float(np.abs(m.stVal[:3] - tr["t_rel"]).max()) for m, tr in zip(maps, truth))
print("worst relative-translation error vs truth %.4f m" % terr)
ntree = min(N, 466)          # longer chains of these maps (distant, low-disparity landmarks kept) are ill conditioned
tree = api.Tree(maps[:ntree])
t = time.time(); tree.solve(); print("merge tree on the first %d built maps: %.4fs" % (ntree, time.time() - t))
s = tree.result_shape(0); print("root m=%d n=%d" % (s.m, s.n))
