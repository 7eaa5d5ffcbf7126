"""The C-ABI library loads and exports every symbol include/linearsfm_b200.h declares; without a
GPU every compute entry fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

from linearsfm_b200 import _lib, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "linearsfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsfm_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_no_cpu_fallback():
    if api.device_count() > 0:
        return
    maps = synth.make_stereo_scene(2, feats_per_frame=8)
    try:
        api.CLinearSFMImp().lmj_PF3D_Divide_ConquerStereo(maps)
    except _lib.LsfmError as e:
        assert e.code == 7
    else:
        raise AssertionError("compute call succeeded without a GPU")


def test_product_does_not_import_oracle():
    # the package AND the developer scripts under tools/: only tests/, smoke() and bench.py's
    # cpu_baseline / reference legs may touch anything under oracle/
    for top in ("linearsfm_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cpp", ".h", ".cuh")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    for word in ("ref_oracle", "oracle/_ref", "libref_", "builder_ref", '"oracle"'):
                        assert word not in txt, (f, word)
