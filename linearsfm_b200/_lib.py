"""ctypes loader of the C-ABI library (include/linearsfm_b200.h). Fails loudly: there is no
Python / CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblinearsfm_b200.so")
CLI_PATH = os.path.join(_HERE, "lib", "LinearSFM")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


class LsfmMap(C.Structure):
    """struct lsfm_map of include/linearsfm_b200.h (mirror of LocalMapInfoStereo/LocalMapInfo)."""
    _fields_ = [("Ref", C.c_int), ("FRef", C.c_int), ("r", C.c_int), ("m", C.c_int), ("n", C.c_int),
                ("nU", C.c_int), ("nW", C.c_int),
                ("ScaP", C.c_int), ("Fix", C.c_int), ("Sign", C.c_int), ("FScaP", C.c_int),
                ("FFix", C.c_int),
                ("stno", _pi), ("stVal", _pd), ("U", _pd), ("Ui", _pi), ("Uj", _pi),
                ("W", _pd), ("photo", _pi), ("feature", _pi), ("V", _pd), ("FBlock", _pi)]


class LsfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lsfm error {code}: {msg}")
        self.code = code


# every symbol include/linearsfm_b200.h declares
EXPORTS = [
    "lsfm_init", "lsfm_shutdown", "lsfm_last_error", "lsfm_device_count", "lsfm_free_map",
    "lsfm_stats_reset", "lsfm_stats_json",
    "lsfm_transform_stereo", "lsfm_transform_stereo_batch", "lsfm_transform_mono", "lsfm_transform_mono_batch", "lsfm_join_stereo",
    "lsfm_join_mono", "lsfm_join_mono_batch", "lsfm_run_mono", "lsfm_run_mono_ex", "lsfm_load_localmap_mono",
    "lsfm_join_stereo_batch", "lsfm_solve_stereo", "lsfm_debug_last_solve", "lsfm_block_ordering",
    "lsfm_run_stereo", "lsfm_tree_create_stereo", "lsfm_tree_solve", "lsfm_tree_result_count",
    "lsfm_tree_result_shape", "lsfm_tree_download", "lsfm_tree_download_state", "lsfm_tree_set_maps",
    "lsfm_tree_free", "lsfm_tree_last_solve_ms", "lsfm_tree_adopt_result", "lsfm_tree_append_maps",
    "lsfm_tree_reset", "lsfm_map_device_bytes", "lsfm_tree_export_device", "lsfm_tree_append_device",
    "lsfm_load_localmap_stereo", "lsfm_save_outputs", "lsfm_cli_main", "lsfm_build_localmaps_stereo", "lsfm_save_localmap",
    "lsfm_pcg_block", "lsfm_marginal_cov_stereo", "lsfm_tree_marginal_cov", "lsfm_solve_mono", "lsfm_save_cache", "lsfm_load_cache", "lsfm_free_cache",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LsfmError(-1, f"{LIB_PATH} is missing: build it with "
                                f"`python -c 'import __graft_entry__ as g; g.build()'` "
                                f"(make -C linearsfm_b200/csrc). There is no fallback path.")
        L = C.CDLL(LIB_PATH)
        L.lsfm_last_error.restype = C.c_char_p
        L.lsfm_stats_json.restype = C.c_char_p
        L.lsfm_free_map.restype = None
        L.lsfm_shutdown.restype = None
        L.lsfm_stats_reset.restype = None
        L.lsfm_tree_free.restype = None
        L.lsfm_tree_free.argtypes = [C.c_void_p]
        L.lsfm_free_cache.restype = None
        L.lsfm_free_cache.argtypes = [C.POINTER(LsfmMap), C.c_int]
        L.lsfm_load_cache.argtypes = [C.c_char_p, C.POINTER(C.POINTER(LsfmMap)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lsfm_save_cache.argtypes = [C.c_char_p, C.POINTER(LsfmMap), C.c_int, C.c_int]
        L.lsfm_tree_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.lsfm_tree_result_count.argtypes = [C.c_void_p]
        L.lsfm_tree_result_shape.argtypes = [C.c_void_p, C.c_int, C.POINTER(LsfmMap)]
        L.lsfm_tree_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(LsfmMap)]
        L.lsfm_tree_download_state.argtypes = [C.c_void_p, C.c_int, _pi, _pd]
        L.lsfm_tree_last_solve_ms.restype = C.c_double
        L.lsfm_tree_last_solve_ms.argtypes = [C.c_void_p]
        L.lsfm_tree_adopt_result.argtypes = [C.c_void_p]
        L.lsfm_tree_reset.argtypes = [C.c_void_p]
        L.lsfm_map_device_bytes.restype = C.c_size_t
        L.lsfm_map_device_bytes.argtypes = [C.POINTER(LsfmMap)]
        L.lsfm_tree_export_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.lsfm_tree_append_device.argtypes = [C.c_void_p, C.POINTER(LsfmMap), C.c_void_p, C.c_size_t]
        L.lsfm_tree_append_maps.argtypes = [C.c_void_p, C.POINTER(LsfmMap), C.c_int]
        L.lsfm_tree_set_maps.argtypes = [C.c_void_p, C.POINTER(LsfmMap), C.c_int]
        L.lsfm_tree_create_stereo.argtypes = [C.POINTER(LsfmMap), C.c_int, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise LsfmError(rc, lib().lsfm_last_error().decode())
