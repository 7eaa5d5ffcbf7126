"""Developer script: run the GPU merge tree at scale with per-stage timing; optional oracle check."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from linearsfm_b200 import api, synth  # noqa: E402

N = int(sys.argv[1]); fpf = int(sys.argv[2]) if len(sys.argv) > 2 else 128
check = (len(sys.argv) > 3 and sys.argv[3] == "check")
t = time.time(); maps = synth.make_stereo_scene(N, feats_per_frame=fpf); print("gen %.2fs" % (time.time() - t))
api.init(0)
t = time.time(); tree = api.Tree(maps); print("upload %.3fs" % (time.time() - t))
for it in range(3):
    api.stats_reset(stage_timing=False)
    t = time.time(); tree.solve(); dt = time.time() - t
    print("solve wall %.4fs launches %d" % (dt, api.stats()["launches"]))
api.stats_reset(stage_timing=True)
t = time.time(); tree.solve(); dt = time.time() - t
st = api.stats()
print("solve (stage timing on) wall %.4fs" % dt)
tot = 0
for k, v in sorted(st["stages"].items()):
    gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
    gfl = v["flops"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
    print("  %-16s %9.3f ms  launches %6d  alg %8.1f MB  %8.1f GB/s %8.2f GF/s" % (k, v["ms"], v["launches"], v["bytes"] / 1e6, gbs, gfl))
    tot += v["ms"]
print("  sum of stages %.3f ms" % tot)
s = tree.result_shape(0); print("root m=%d n=%d nU=%d nW=%d" % (s.m, s.n, s.nU, s.nW))
if check:
    import ref_oracle as ro
    got = tree.download(0)
    ref, tref, twall = ro.run_tree_stereo(maps)
    print("oracle: ref clock %.3fs wall %.3fs shim %s" % (tref, twall, ro.shim_times()))
    from linearsfm_b200.localmap import maps_equal_int
    print("int mismatches:", maps_equal_int(got, ref))
    def rel(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300))
    print("rel err stVal %.3e U %.3e W %.3e V %.3e" % (rel(got.stVal, ref.stVal), rel(got.U, ref.U), rel(got.W, ref.W), rel(got.V, ref.V)))
