"""Developer script: mono merge tree at RS90 / RS468 sizes, GPU vs oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from linearsfm_b200 import api, synth
from linearsfm_b200.localmap import maps_equal_int
import ref_oracle as ro
N = int(sys.argv[1]); fpf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
t = time.time(); maps = synth.make_mono_scene(N, feats_per_frame=fpf); print("gen %.1fs, mean n %.0f" % (time.time() - t, np.mean([m.n for m in maps])))
api.init(0)
for it in range(3):
    t = time.perf_counter(); got = api.run_mono(maps); dt = time.perf_counter() - t
    print("gpu run_mono (host buffers in/out) %.4f s" % dt)
ref, tref, twall = ro.run_tree_mono(maps)
print("oracle %.3f s (clock) %.3f s (wall); root m=%d n=%d nU=%d nW=%d" % (tref, twall, ref.m, ref.n, ref.nU, ref.nW))
print("int mismatches:", maps_equal_int(got, ref), "meta", [(getattr(got, k), getattr(ref, k)) for k in ("Ref", "ScaP", "Fix", "Sign")])
rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300))
print("rel err stVal %.3e U %.3e W %.3e V %.3e" % (rel(got.stVal, ref.stVal), rel(got.U, ref.U), rel(got.W, ref.W), rel(got.V, ref.V)))
