"""Developer script: where does the host-buffer (e2e) path spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from linearsfm_b200 import api, synth, dist as lsd
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3499
maps = synth.make_stereo_scene(N, feats_per_frame=128)
api.init(0)
arr, keep = api.to_c_array(maps)
be = lsd.TreeBackend(api, maps)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    be.tree.set_maps_c(arr, len(maps)); torch.cuda.synchronize(); t1 = time.perf_counter()
    be.tree.solve(); torch.cuda.synchronize(); t2 = time.perf_counter()
    be.tree.download_state(0); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("upload %.2f ms  solve %.2f ms  download_state %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
# raw PCIe numbers for comparison
h = torch.empty(356 * 1000 * 1000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print("raw H2D 356 MB: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
