"""Multi-GPU parity: the tree sharded over 2 (or 4) GPUs with NCCL hand-over equals the oracle.
Skipped on single-GPU boxes (the schedule itself is covered on CPU by test_dist_gloo.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from linearsfm_b200 import synth
from util import rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,n", [(2, 13), (2, 32), (4, 27)])
def test_sharded_gpu_tree(gpu, oracle, tmp_path, world, n):
    if gpu.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "fin.npz")
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_gpu_worker.py"), str(n), out], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    got = np.load(out)
    maps = synth.make_stereo_scene(n, feats_per_frame=24, seed=500 + n)
    ref, _, _ = oracle.run_tree_stereo(maps)
    assert list(got["meta"]) == [ref.Ref, ref.FRef, ref.m, ref.n]
    for f in ("stno", "Ui", "Uj", "photo", "feature", "FBlock"):
        assert np.array_equal(got[f], getattr(ref, f)), f
    assert rel_err(got["stVal"], ref.stVal) <= 1e-6
    for f in ("U", "W", "V"):
        assert rel_err(got[f].reshape(-1), getattr(ref, f).reshape(-1)) <= 1e-7, f
