"""Stress run: synthetic NC3500-shape scene scaled up (BASELINE.json configs[4]: 50k local maps).
The reference cannot run this size (int overflow / O(m^2) mask, LinearSFMImp.cpp:2131).  The scene is the
well-conditioned variant (the trajectory closes on itself: four laps by default, outlier-gated landmarks):
a 50k-frame OPEN chain is numerically hopeless for any FP64 solver.  Two solves: results must be
bit-identical.
Usage: python tools/stress.py N landmarks_per_frame [lap_frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from linearsfm_b200 import api, synth
N = int(sys.argv[1]); fpf = int(sys.argv[2])
t = time.time()
lap = int(sys.argv[3]) if len(sys.argv) > 3 else max(500, N // 4)
maps, truth = synth.make_stereo_scene(N, feats_per_frame=fpf, return_truth=True, revisit=0.1, lap=lap, max_depth=15.0, gate=True)
print("gen %.1fs, %d landmarks rows" % (time.time() - t, sum(m.n for m in maps)), flush=True)
api.init(0)
t = time.time(); tree = api.Tree(maps); print("upload %.2fs" % (time.time() - t), flush=True)
states = []
for it in range(2):
    tree.reset()
    t = time.perf_counter(); tree.solve(); dt = time.perf_counter() - t
    s = tree.result_shape(0)
    print("solve %d wall %.3f s device %.1f ms  root m=%d n=%d nU=%d nW=%d" % (it, dt, tree.last_solve_ms(), s.m, s.n, s.nU, s.nW), flush=True)
    states.append(tree.download_state(0))
print("two solves bit-identical:", bool(np.array_equal(states[0][0], states[1][0]) and np.array_equal(states[0][1], states[1][1])))
stno, st = states[0]
m = s.m
P = st[:6 * m].reshape(m, 6); pid = -stno[:6 * m:6]
print("pose error vs truth (max |dt|): %.3f m over %d frames, all finite: %s" % (np.abs(P[:, :3] - truth["pose_t"][pid - 1]).max(), N, bool(np.all(np.isfinite(st)))))
api.stats_reset(stage_timing=True); tree.reset(); tree.solve(); st = api.stats()
for k, v in sorted(st["stages"].items()):
    print("  %-18s %9.2f ms  %8.1f GB/s" % (k, v["ms"], v["bytes"] / max(v["ms"], 1e-9) / 1e6))
