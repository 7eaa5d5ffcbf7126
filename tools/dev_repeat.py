"""Developer script: repeated resident solves, per-solve wall + device ms (variance check)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linearsfm_b200 import api, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3499
maps = synth.make_stereo_scene(N, feats_per_frame=128)
api.init(0)
tree = api.Tree(maps)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    t = time.perf_counter(); tree.solve(); dt = time.perf_counter() - t
    print("solve %2d wall %.4f s device %.2f ms" % (it, dt, tree.last_solve_ms()), flush=True)
arr, keep = api.to_c_array(maps)
for it in range(4):
    t = time.perf_counter(); tree.set_maps_c(arr, len(maps)); t1 = time.perf_counter(); tree.solve(); t2 = time.perf_counter(); tree.download_state(0); t3 = time.perf_counter()
    print("e2e %d upload %.4f solve %.4f (device %.2f ms) download %.4f" % (it, t1 - t, t2 - t1, tree.last_solve_ms(), t3 - t2), flush=True)
