"""Shared helpers for the test-suite."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "emu_render.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libm2h_emu.so")


def rel_l2(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def f16_ulp_diff(a, b):
    """|a - b| in units of fp16 ulps (monotone integer mapping of the bit patterns)."""
    def key(x):
        u = np.ascontiguousarray(x, dtype=np.float16).view(np.uint16).astype(np.int32)
        return np.where(u & 0x8000, 0x8000 - u, u)
    return np.abs(key(a) - key(b))


def load_emu():
    """Build (g++) and load the CPU emulation of the render kernels - test infrastructure only."""
    csrc = os.path.join(ROOT, "move2hear_active_av_separation_b200", "csrc")
    deps = [EMU_SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", EMU_LIB, EMU_SRC])
    lib = ctypes.CDLL(EMU_LIB)
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# fp16 end-to-end tolerance: 1 fp16 ulp of the output, or 5e-4 absolute.  The observation is
# fp16(log1p(fp16(|STFT|))): a 1-LSB flip of the quantised waveform (which already separates two exact CPU routes,
# tests/test_oracle.py) can move the fp16 magnitude by one ulp (4.9e-4 in [0.5, 1)), i.e. two ulps of the
# log1p output in [0.25, 0.5), and several ulps of near-zero bins whose ulp is < 1e-5.
def f16_close(a, b, abs_tol=5e-4, max_bad_frac=0.0):
    ulp = f16_ulp_diff(a, b)
    bad = (ulp > 1) & (np.abs(a.astype(np.float32) - b.astype(np.float32)) > abs_tol)
    return bad.mean() <= max_bad_frac, int(bad.sum()), int(ulp.max())
