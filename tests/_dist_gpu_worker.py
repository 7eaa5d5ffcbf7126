"""Worker of test_gpu_multi.py: one rank per GPU, NCCL, sharded merge tree; rank 0 saves the result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from linearsfm_b200 import api, dist as lsd, synth  # noqa: E402


def main():
    n, out = int(sys.argv[1]), sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.init(local)
    maps = synth.make_stereo_scene(n, feats_per_frame=24, seed=500 + n)
    lo, hi = lsd.slice_of(n, world, rank)
    be = lsd.TreeBackend(api, maps[lo:hi])
    root = lsd.run_sharded(be, n, rank, world, torch.device("cuda", local))
    if root:
        fin = be.tree.download(0)
        np.savez(out, stno=fin.stno, stVal=fin.stVal, U=fin.U, Ui=fin.Ui, Uj=fin.Uj, W=fin.W,
                 photo=fin.photo, feature=fin.feature, V=fin.V, FBlock=fin.FBlock,
                 meta=np.array([fin.Ref, fin.FRef, fin.m, fin.n]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
