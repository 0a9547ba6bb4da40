"""Build + load ``libm2h_b200.so`` (the C-ABI library declared in ``include/m2h.h``).

The library is compiled in-tree with nvcc for sm_100a only and loaded with ctypes; there is no
CPU fallback: every compute entry point needs a CUDA device and raises :class:`M2HError` otherwise.
"""
from __future__ import annotations

import ctypes
import fcntl
import glob
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libm2h_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


class M2HError(RuntimeError):
    pass


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise M2HError("nvcc not found - cannot build libm2h_b200.so")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = glob.glob(os.path.join(CSRC, "*")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` into ``libm2h_b200.so`` (sm_100a, -lineinfo).

    Many processes may import the package at once (8 ranks x 14 env workers, DataLoader workers): the build runs under
    an inter-process file lock and re-checks staleness once it holds the lock, so exactly one process compiles.
    """
    if not force and not _stale():
        return LIB_PATH
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    with open(os.path.join(build_dir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another process built it while we waited
                return LIB_PATH
            return _build_locked(build_dir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(build_dir: str, verbose: bool) -> str:
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise M2HError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    # libcuda is only needed at run time (driver entry points are resolved through cudart)
    cmd = [_nvcc(), "-shared", "-o", tmp] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise M2HError("link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lock = threading.Lock()
_lib = None

_vp, _i, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/m2h.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "m2h_abi_version": (_i, []),
    "m2h_last_error": (ctypes.c_char_p, []),
    "m2h_init": (_i, [_i]),
    "m2h_render_workspace_bytes": (_sz, [_i]),
    "m2h_rir_convolve": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "m2h_rir_convolve_indexed": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "m2h_wav_parse": (_i, [_vp, _sz, _vp]),
    "m2h_rir_ingest": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "m2h_clip_ingest": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "m2h_stft_waves": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "m2h_render_mixture": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "m2h_stft_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "m2h_istft": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "m2h_unet_enc_packed_bytes": (_sz, [_i]),
    "m2h_unet_dec_packed_bytes": (_sz, [_i]),
    "m2h_unet_pack_encoder": (_i, [_vp, _i, ctypes.c_float, _vp, _vp]),
    "m2h_unet_pack_decoder": (_i, [_vp, _i, ctypes.c_float, _vp, _vp]),
    "m2h_unet_workspace_bytes": (_sz, [_i]),
    "m2h_unet_workspace_layout": (_i, [_i, _vp]),
    "m2h_unet_workspace_clear_halos": (_i, [_i, _vp, _sz, _vp]),
    "m2h_unet_encoder_forward": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_decoder_forward": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_train_workspace_bytes": (_sz, [_i]),
    "m2h_unet_train_workspace_layout": (_i, [_i, _vp]),
    "m2h_unet_train_workspace_init": (_i, [_i, _vp, _sz, _vp]),
    "m2h_stft_waves_x0": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "m2h_unet_encoder_forward_x0": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_pack_train_weights": (_i, [_vp, _vp, _i, _i, _vp, _sz, _vp]),
    "m2h_unet_encoder_forward_train": (_i, [_vp, _vp, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_decoder_forward_train": (_i, [_vp, _vp, _i, ctypes.c_float, ctypes.c_float, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_decoder_backward": (_i, [_vp, _i, _vp, _i, _vp, _sz, _vp, _vp]),
    "m2h_unet_encoder_backward": (_i, [_vp, _i, _vp, _i, _vp, _sz, _vp, _vp]),
    "m2h_l1_mask_loss": (_i, [_vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _vp, _vp, _vp]),
    "m2h_stft_l2_distance": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "m2h_wave_moments": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "m2h_acoustic_mem_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, _i, _vp, _vp]),
}


def load(build_if_missing: bool = True):
    """Return the ctypes handle (building the library first if it is missing or stale)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing and _stale() and os.path.exists(_nvcc_or_none() or ""):
            build()
        if not os.path.exists(LIB_PATH):
            raise M2HError("libm2h_b200.so is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = ABI drift, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def _nvcc_or_none():
    try:
        return _nvcc()
    except M2HError:
        return None


def check(rc: int, what: str = "m2h"):
    if rc != 0:
        msg = load().m2h_last_error()
        raise M2HError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
