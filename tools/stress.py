"""Stress run: synthetic NC3500-shape scene scaled up (BASELINE.json configs[4])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from linearsfm_b200 import api, synth
N = int(sys.argv[1]); fpf = int(sys.argv[2])
t = time.time(); maps, truth = synth.make_stereo_scene(N, feats_per_frame=fpf, return_truth=True); print("gen %.1fs" % (time.time() - t), flush=True)
api.init(0)
t = time.time(); tree = api.Tree(maps); print("upload %.2fs" % (time.time() - t), flush=True)
for it in range(2):
    t = time.perf_counter(); tree.solve(); dt = time.perf_counter() - t
    s = tree.result_shape(0)
    print("solve %d wall %.3f s device %.1f ms  root m=%d n=%d nU=%d nW=%d" % (it, dt, tree.last_solve_ms(), s.m, s.n, s.nU, s.nW), flush=True)
stno, st = tree.download_state(0)
m = s.m
P = st[:6 * m].reshape(m, 6); pid = -stno[:6 * m:6]
print("pose drift vs truth (max |dt|): %.3f m over %d frames" % (np.abs(P[:, :3] - truth["pose_t"][pid - 1]).max(), N))
api.stats_reset(stage_timing=True); tree.solve(); st = api.stats()
for k, v in sorted(st["stages"].items()):
    print("  %-18s %9.2f ms  %8.1f GB/s" % (k, v["ms"], v["bytes"] / max(v["ms"], 1e-9) / 1e6))
