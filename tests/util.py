import numpy as np

from linearsfm_b200.localmap import LocalMap, maps_equal_int


def rel_err(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    if a.size == 0 and b.size == 0:
        return 0.0
    if a.shape != b.shape:
        return float("inf")
    s = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def assert_maps_match(got: LocalMap, ref: LocalMap, tol_state=1e-9, tol_info=1e-9, what=""):
    bad = maps_equal_int(got, ref)
    assert not bad, f"{what}: integer fields differ: {bad}"
    e = rel_err(got.stVal, ref.stVal)
    assert e <= tol_state, f"{what}: stVal rel err {e:.3e}"
    for name in ("U", "W", "V"):
        e = rel_err(getattr(got, name), getattr(ref, name))
        assert e <= tol_info, f"{what}: {name} rel err {e:.3e}"


def state_rel_err(got: LocalMap, ref: LocalMap):
    """north_star tolerance: poses and features agree within 1e-6 relative."""
    assert np.array_equal(got.stno, ref.stno)
    return rel_err(got.stVal, ref.stVal)
