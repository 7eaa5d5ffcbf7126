"""Developer check: what a rank without leaf maps does in bench.py (3499 maps on 8 GPUs leave rank 7 empty)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linearsfm_b200 import api, dist as lsd
api.init(0)
arr, keep = api.to_c_array([])
be = lsd.TreeBackend(api, [])
be.tree.reset()
be.tree.set_maps_c(arr, 0)
print("empty rank ok: count", be.count(), "last_solve_ms", be.tree.last_solve_ms(), "launches", api.stats()["launches"])
lo, hi = lsd.slice_of(3499, 8, 7); print("slice of rank 7:", lo, hi, "plan", lsd.plan(3499, 8)[:2])
