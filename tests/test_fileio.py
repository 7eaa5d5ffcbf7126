"""Drop-in file formats: our C++ reader/writers vs the reference's own (through the oracle)."""
import ctypes as C
import filecmp
import os

import numpy as np

from linearsfm_b200 import _lib, api, synth
from linearsfm_b200.localmap import write_localmap, read_localmap
from util import assert_maps_match


def test_reader_matches_reference(oracle, tmp_path):
    maps = synth.make_stereo_scene(3, feats_per_frame=10, seed=5)
    for i, lm in enumerate(maps):
        p = str(tmp_path / f"localmap_{i + 1}.txt")
        write_localmap(p, lm)
        ref = oracle.load_localmap_stereo(p)
        out = _lib.LsfmMap()
        _lib.check(_lib.lib().lsfm_load_localmap_stereo(p.encode(), C.byref(out)))
        got = api.from_c(out)
        assert_maps_match(got, ref, tol_state=0, tol_info=0, what="reader")
        assert_maps_match(got, lm, tol_state=0, tol_info=0, what="round trip")
        assert_maps_match(read_localmap(p), lm, tol_state=0, tol_info=0, what="python reader")


def test_reader_errors(tmp_path):
    out = _lib.LsfmMap()
    assert _lib.lib().lsfm_load_localmap_stereo(str(tmp_path / "nope.txt").encode(), C.byref(out)) == 5
    p = tmp_path / "bad.txt"
    p.write_text("1\n9\n-2 0.5\n")
    assert _lib.lib().lsfm_load_localmap_stereo(str(p).encode(), C.byref(out)) == 6


def test_writers_match_reference(oracle, tmp_path):
    maps = synth.make_stereo_scene(4, feats_per_frame=10, seed=6)
    fin, _, _ = oracle.run_tree_stereo(maps)
    oracle.save_outputs(fin, st=str(tmp_path / "st_ref.txt"), pose=str(tmp_path / "p_ref.txt"),
                        feat=str(tmp_path / "f_ref.txt"))
    c, _keep = api.to_c(fin)
    rc = _lib.lib().lsfm_save_outputs(C.byref(c), str(tmp_path / "st.txt").encode(),
                                      str(tmp_path / "p.txt").encode(), str(tmp_path / "f.txt").encode())
    assert rc == 0
    for a in ("st", "p", "f"):
        assert filecmp.cmp(tmp_path / f"{a}.txt", tmp_path / f"{a}_ref.txt", shallow=False), a


def test_mono_reader_matches_reference(oracle, tmp_path):
    import ctypes as C
    maps = synth.make_mono_scene(3, feats_per_frame=10, seed=8)
    for i, lm in enumerate(maps):
        p = str(tmp_path / f"localmap_{i + 1}.txt")
        write_localmap(p, lm, mono=True)
        out = oracle.RefMap()
        assert oracle.lib().ref_load_localmap_mono(p.encode(), C.byref(out)) == 0
        ref = oracle.from_c(out, free=False)
        got_c = _lib.LsfmMap()
        _lib.check(_lib.lib().lsfm_load_localmap_mono(p.encode(), C.byref(got_c)))
        got = api.from_c(got_c)
        assert_maps_match(got, ref, tol_state=0, tol_info=0, what="mono reader")
        for k in ("ScaP", "Fix", "Sign", "FScaP", "FFix"):
            assert getattr(got, k) == getattr(ref, k) == getattr(lm, k), k


def test_localmap_writer_is_read_back_by_the_reference(oracle, tmp_path):
    # lsfm_save_localmap (the joined map with its information matrix as an output):
    # the REFERENCE's own reader (fscanf) and ours get every number back bit for bit
    maps = synth.make_stereo_scene(4, feats_per_frame=12, seed=6)
    joined, _, _ = oracle.run_tree_stereo(maps)            # a map with several poses and W fill-in
    for k, lm in enumerate([maps[0], joined]):
        p = str(tmp_path / f"out_{k}.txt")
        c, keep = api.to_c(lm)
        _lib.check(_lib.lib().lsfm_save_localmap(C.byref(c), p.encode(), C.c_int(0)))
        ref = oracle.load_localmap_stereo(p)
        assert_maps_match(ref, lm, tol_state=0, tol_info=0, what="reference reads our localmap file")
        out = _lib.LsfmMap()
        _lib.check(_lib.lib().lsfm_load_localmap_stereo(p.encode(), C.byref(out)))
        assert_maps_match(api.from_c(out), lm, tol_state=0, tol_info=0, what="our reader, our writer")
    assert _lib.lib().lsfm_save_localmap(C.byref(c), str(tmp_path / "no_dir" / "x.txt").encode(), C.c_int(0)) == 5


def test_localmap_writer_parallel_path(oracle, tmp_path):
    # above 50,000 elements per array lsfm_save_localmap formats on several host threads, a wave at a time:
    # the file is what one fprintf("%.17g") per element would write (Python's "%.17g" is the same conversion),
    # and the reference's reader gets every number back bit for bit (values across 16 decades, negative zero)
    from linearsfm_b200.localmap import LocalMap
    m, n, nW = 40, 6000, 9000
    rng = np.random.default_rng(3)
    stno = np.concatenate([np.repeat(-np.arange(2, m + 2), 6), np.repeat(np.arange(n) + 1, 3)]).astype(np.int32)
    W = rng.normal(size=(nW, 6, 3)) * 10.0 ** rng.integers(-8, 8, (nW, 1, 1))
    W[5, 0, 0] = -0.0
    feature = np.sort(rng.integers(0, n, nW)).astype(np.int32)
    FBlock = np.full(n, -1, np.int32)
    first = np.flatnonzero(np.r_[True, feature[1:] != feature[:-1]])
    FBlock[feature[first]] = first
    lm = LocalMap(Ref=3, stno=stno, stVal=rng.normal(0, 50, 6 * m + 3 * n), m=m, n=n, U=rng.normal(size=(m, 6, 6)),
                  Ui=np.arange(m), Uj=np.arange(m), W=W, photo=rng.integers(0, m, nW), feature=feature,
                  V=rng.normal(size=(n, 3, 3)), FBlock=FBlock)
    p = str(tmp_path / "big.txt")
    c, keep = api.to_c(lm)
    _lib.check(_lib.lib().lsfm_save_localmap(C.byref(c), p.encode(), C.c_int(0)))
    assert_maps_match(oracle.load_localmap_stereo(p), lm, tol_state=0, tol_info=0, what="reference reads the big localmap file")

    def rows(x, per):
        x = np.asarray(x, np.float64).reshape(-1)
        return "".join("%.17g%s" % (v, "\n" if (i + 1) % per == 0 or i + 1 == x.size else " ") for i, v in enumerate(x))

    def ints(x):
        return " ".join(str(int(v)) for v in x) + "\n"
    want = "%d\n%d\n" % (lm.Ref, lm.r) + "".join("%d %.17g\n" % (a, v) for a, v in zip(lm.stno, lm.stVal))
    want += "%d %d\n%d\n" % (m, n, m) + rows(lm.U, 6) + ints(lm.Ui) + ints(lm.Uj) + "%d\n" % nW + rows(lm.W, 3)
    want += ints(lm.photo) + ints(lm.feature) + rows(lm.V, 3) + ints(lm.FBlock)
    assert open(p).read() == want


def test_parallel_writers_byte_identical_on_a_large_map(oracle, tmp_path):
    # the multi-threaded formatting path (>= 20000 rows) against the reference's fprintf loops: same bytes,
    # including a repeated landmark id (the last occurrence wins, std::map semantics of 7907/7925),
    # a huge value, a negative zero and a value that rounds to 0.000001
    from linearsfm_b200.localmap import LocalMap
    m, n = 300, 40000
    rng = np.random.default_rng(0)
    ids = rng.permutation(n) + 1
    ids[1000] = ids[5]
    stno = np.concatenate([np.repeat(-np.arange(2, m + 2), 6), np.repeat(ids, 3)]).astype(np.int32)
    stVal = rng.normal(0, 50, 6 * m + 3 * n)
    stVal[7], stVal[8], stVal[9] = 1e300, -0.0, 5e-7
    lm = LocalMap(Ref=1, stno=stno, stVal=stVal, m=m, n=n, U=np.zeros((1, 6, 6)), Ui=[0], Uj=[0],
                  W=np.zeros((1, 6, 3)), photo=[0], feature=[0], V=np.zeros((n, 3, 3)), FBlock=np.zeros(n, np.int32))
    c, keep = api.to_c(lm)
    got = {k: str(tmp_path / f"got_{k}.txt") for k in ("st", "p", "f")}
    ref = {k: str(tmp_path / f"ref_{k}.txt") for k in ("st", "p", "f")}
    _lib.check(_lib.lib().lsfm_save_outputs(C.byref(c), got["st"].encode(), got["p"].encode(), got["f"].encode()))
    oracle.save_outputs(lm, st=ref["st"], pose=ref["p"], feat=ref["f"])
    for k in got:
        assert filecmp.cmp(got[k], ref[k], shallow=False), k


def test_binary_cache_round_trip(tmp_path):
    """lsfm_save_cache / lsfm_load_cache: the maps come back bit for bit; a truncated file and a
    file of another kind are refused."""
    from linearsfm_b200.localmap import maps_equal_int
    maps = synth.make_stereo_scene(5, feats_per_frame=12, seed=11)
    arr, _keep = api.to_c_array(maps)
    f = str(tmp_path / "maps.lsfmcache")
    _lib.check(_lib.lib().lsfm_save_cache(f.encode(), arr, C.c_int(len(maps)), C.c_int(0)))
    out = C.POINTER(_lib.LsfmMap)()
    num, mono = C.c_int(0), C.c_int(-1)
    _lib.check(_lib.lib().lsfm_load_cache(f.encode(), C.byref(out), C.byref(num), C.byref(mono)))
    assert num.value == len(maps) and mono.value == 0
    for i, lm in enumerate(maps):
        got = api.from_c(out[i], free=False)
        assert not maps_equal_int(got, lm)
        for name in ("stVal", "U", "W", "V"):
            assert np.array_equal(getattr(got, name), getattr(lm, name)), name
    _lib.lib().lsfm_free_cache(out, num)
    raw = open(f, "rb").read()
    open(f, "wb").write(raw[: len(raw) // 2])
    assert _lib.lib().lsfm_load_cache(f.encode(), C.byref(out), C.byref(num), C.byref(mono)) == 6
    open(f, "wb").write(b"not a cache at all")
    assert _lib.lib().lsfm_load_cache(f.encode(), C.byref(out), C.byref(num), C.byref(mono)) == 6
    assert _lib.lib().lsfm_load_cache(str(tmp_path / "missing").encode(), C.byref(out), C.byref(num), C.byref(mono)) == 5
