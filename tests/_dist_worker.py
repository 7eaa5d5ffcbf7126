"""Worker of test_dist_gloo.py: one rank of the sharded merge tree over gloo, CPU only.
The compute backend is the oracle (reference operators), so the test isolates the HOST logic of
linearsfm_b200/dist.py: slicing, global-index re-base rule, hand-over schedule, (de)serialisation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch.distributed as dist  # noqa: E402
import ref_oracle as ro  # noqa: E402
from linearsfm_b200 import dist as lsd, synth  # noqa: E402


class OracleBackend:
    def __init__(self, maps):
        self.maps = list(maps)

    def solve_levels(self, first_index, levels):
        base = first_index
        maps = self.maps
        for _ in range(levels):
            if not maps:
                break
            nxt = []
            for i in range(len(maps) // 2):
                E, C = maps[2 * i], maps[2 * i + 1]
                Et = E if E.Ref == C.Ref else ro.transform_stereo(E, C.Ref)
                nxt.append(ro.join_stereo(Et, C))
            if len(maps) % 2:
                nxt.append(maps[-1])
            base //= 2
            for i, mp in enumerate(nxt):
                if (base + i + 1) % 2 == 0 and mp.Ref > mp.FRef:
                    nxt[i] = ro.transform_stereo(mp, mp.FRef)
            maps = nxt
        self.maps = maps

    def count(self):
        return len(self.maps)

    def get(self, i):
        return self.maps[i]

    def set(self, maps):
        self.maps = list(maps)

    def append(self, maps):
        self.maps += list(maps)

    def finish(self):
        mp = self.maps[0]
        if mp.Ref > mp.FRef:
            self.maps[0] = ro.transform_stereo(mp, mp.FRef)


def main():
    n, out = int(sys.argv[1]), sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    maps = synth.make_stereo_scene(n, feats_per_frame=10, seed=300 + n)
    lo, hi = lsd.slice_of(n, world, rank)
    be = OracleBackend(maps[lo:hi])
    is_root = lsd.run_sharded(be, n, rank, world)
    if is_root:
        fin = be.get(0)
        np.savez(out, stno=fin.stno, stVal=fin.stVal, U=fin.U, Ui=fin.Ui, Uj=fin.Uj, W=fin.W,
                 photo=fin.photo, feature=fin.feature, V=fin.V, FBlock=fin.FBlock,
                 meta=np.array([fin.Ref, fin.FRef, fin.m, fin.n]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
