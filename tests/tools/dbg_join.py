"""Debug: N=2 tree stage by stage against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from linearsfm_b200 import api as gpu, synth
import ref_oracle as oracle
from util import rel_err
n, fpf = int(sys.argv[1]), int(sys.argv[2])
maps = synth.make_stereo_scene(n, feats_per_frame=fpf, seed=100 + n)
gpu.init(0)
for rep in range(2):
    e = oracle.transform_stereo(maps[0], maps[1].Ref)
    ref = oracle.join_stereo(e, maps[1])
    got = gpu.join_stereo_batch([e], [maps[1]])[0]
    print("join: m n nW", ref.m, ref.n, len(ref.photo), "stVal", rel_err(got.stVal, ref.stVal), "W", rel_err(got.W, ref.W),
          "V", rel_err(got.V, ref.V), "U", rel_err(got.U, ref.U), "photo eq", np.array_equal(got.photo, ref.photo))
    d = np.abs(np.asarray(got.stVal) - np.asarray(ref.stVal))
    print("  worst idx", int(np.argmax(d)), "of", len(d), "stno", ref.stno[int(np.argmax(d))], "got", got.stVal[int(np.argmax(d))], "ref", ref.stVal[int(np.argmax(d))])
    reft, _, _ = oracle.run_tree_stereo(maps)
    gott = gpu.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(maps)
    print("tree: stVal", rel_err(gott.stVal, reft.stVal), "W", rel_err(gott.W, reft.W))
