"""N>1 path on CPU: the sharded merge-tree schedule (linearsfm_b200/dist.py) over gloo with
world_size 2 and 4 reproduces the reference's sequential scheduler bit for bit when both use the
same operators (the oracle's)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from linearsfm_b200 import dist as lsd, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_plan():
    assert lsd.plan(3499, 8) == (512, 9, list(range(7)))
    assert lsd.plan(16, 2) == (8, 3, [0, 1])
    assert lsd.plan(5, 2) == (4, 2, [0, 1])
    assert lsd.plan(3, 4) == (1, 0, [0, 1, 2])
    assert lsd.slice_of(3499, 8, 6) == (3072, 3499) and lsd.slice_of(3499, 8, 7) == (3499, 3499)


def test_pack_roundtrip():
    lm = synth.make_stereo_scene(2, feats_per_frame=6)[1]
    got = lsd.unpack_map(*lsd.pack_map(lm))
    for f in ("stno", "stVal", "U", "Ui", "Uj", "W", "photo", "feature", "V", "FBlock"):
        assert np.array_equal(getattr(got, f), getattr(lm, f))
    assert (got.Ref, got.FRef, got.m, got.n) == (lm.Ref, lm.FRef, lm.m, lm.n)


@pytest.mark.parametrize("world,n", [(2, 5), (2, 8), (2, 11), (4, 7), (4, 13), (4, 5), (8, 27)])   # the last two leave a rank without maps (like 3499 maps on 8 GPUs)
def test_sharded_schedule_matches_sequential(oracle, tmp_path, world, n):
    out = str(tmp_path / "fin.npz")
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py"), str(n), out], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(out)
    maps = synth.make_stereo_scene(n, feats_per_frame=10, seed=300 + n)
    ref, _, _ = oracle.run_tree_stereo(maps)
    assert list(got["meta"]) == [ref.Ref, ref.FRef, ref.m, ref.n]
    for f in ("stno", "Ui", "Uj", "photo", "feature", "FBlock", "stVal", "U", "W", "V"):
        assert np.array_equal(got[f], getattr(ref, f)), f
