"""CPU: the C-ABI library loads and exports every symbol include/m2h.h declares (no compute calls)."""
import os
import re

import pytest

from move2hear_active_av_separation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "m2h.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m2h_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for " + n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.m2h_abi_version() == 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from move2hear_active_av_separation_b200.render import AudioRenderer
    with pytest.raises(_lib.M2HError):
        AudioRenderer()
    # the raw ABI fails loudly too (CUDA error, not a silent CPU path)
    lib = _lib.load()
    assert lib.m2h_init(0) != 0 and lib.m2h_last_error()
