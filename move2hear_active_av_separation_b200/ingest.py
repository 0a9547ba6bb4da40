"""wav ingest + device-side caches of the clips and RIRs.

The reference reads two kinds of wav files on the per-step path:

* the monaural clip of each source, cached on the host by sound name
  (``_cached_mono_waveforms``, habitat_audio/simulator_train.py:396-400; ``file2audio_dict``, dataset.py:155-160);
* the binaural RIR of each (scene, azimuth, receiver, source), re-read from disk every step
  (simulator_train.py:402-415; dataset.py:168-176), with zeros substituted for unreadable / empty files.

Here both live in device-resident pools with LRU replacement, so that a step whose clips and RIRs were seen before
moves only two small index arrays across PCIe (the render is ~1 ms per 1024 mixtures; the RIRs alone are 256 KB per
mixture).  A miss is decoded by :func:`parse_wav` (host: RIFF chunk walk, ``m2h_wav_parse``), staged in pinned memory,
copied once per batch and turned into a pool row on the device (``m2h_rir_ingest`` / ``m2h_clip_ingest``: de-interleave,
cast, zero tail).  The convolution then reads its inputs through the index arrays (``m2h_rir_convolve_indexed``).
There is no CPU fallback: every entry point needs the CUDA library and a device.
"""
from __future__ import annotations

import ctypes
import logging
from collections import OrderedDict
from typing import Callable, Hashable, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .render import AudioRenderer, _stream_ptr

WAV_PCM16, WAV_PCM32, WAV_FLOAT32 = 1, 2, 3
_NP_FORMAT = {np.dtype(np.int16): WAV_PCM16, np.dtype(np.int32): WAV_PCM32, np.dtype(np.float32): WAV_FLOAT32}
_FORMAT_BYTES = {WAV_PCM16: 2, WAV_PCM32: 4, WAV_FLOAT32: 4}


class WavInfo(ctypes.Structure):
    """Mirror of ``m2h_wav_info`` (include/m2h.h)."""
    _fields_ = [("format", ctypes.c_int), ("channels", ctypes.c_int), ("sample_rate", ctypes.c_int),
                ("bits_per_sample", ctypes.c_int), ("frames", ctypes.c_longlong), ("data_offset", ctypes.c_size_t)]


def parse_wav(data: Union[bytes, bytearray, memoryview, np.ndarray]) -> WavInfo:
    """Locate the samples of a RIFF/WAVE image (host only; what ``scipy.io.wavfile.read`` does before it builds the
    array).  Raises :class:`ValueError` for anything that is not PCM16 / PCM32 / float32 - the reference treats a
    ``ValueError`` from ``wavfile.read`` as "file is not readable" (simulator_train.py:411-413)."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    info = WavInfo()
    lib = _lib.load()
    rc = lib.m2h_wav_parse(buf.ctypes.data_as(ctypes.c_void_p), buf.size, ctypes.byref(info))
    if rc != 0:
        raise ValueError(lib.m2h_last_error().decode())
    return info


class RowCache:
    """LRU map ``key -> row`` of a fixed-capacity pool.  Rows touched by the batch being assembled are pinned."""

    def __init__(self, capacity: int):
        if capacity < 1:
            raise ValueError("capacity must be positive")
        self.capacity = capacity
        self.rows: "OrderedDict[Hashable, int]" = OrderedDict()
        self.free = list(range(capacity - 1, -1, -1))
        self.hits = 0
        self.misses = 0

    def lookup(self, key: Hashable, pinned: set):
        """Returns ``(row, hit)``; on a miss the row is the one to fill (least recently used, not in ``pinned``)."""
        row = self.rows.get(key)
        if row is not None:
            self.rows.move_to_end(key)
            self.hits += 1
            pinned.add(row)
            return row, True
        self.misses += 1
        if self.free:
            row = self.free.pop()
        else:
            victim = next((k for k, r in self.rows.items() if r not in pinned), None)
            if victim is None:
                raise _lib.M2HError(f"cache of {self.capacity} rows is smaller than one batch's distinct keys")
            row = self.rows.pop(victim)
        self.rows[key] = row
        pinned.add(row)
        return row, False

    def forget(self, key: Hashable):
        """Undo a miss whose row was never filled: the key is unmapped and its row returns to the free list."""
        row = self.rows.pop(key, None)
        if row is not None:
            self.free.append(row)


RirSource = Union[None, bytes, bytearray, memoryview, np.ndarray]


class CachedAudioRenderer:
    """Render mixtures from *keys*: clips by sound name, RIRs by (scene, azimuth, receiver, source).

    ``clip_loader(name)`` / ``rir_loader(key)`` are called for misses only and return either the wav file image
    (``bytes``) or an array (clip ``int16 [N]``; RIR ``[L, 2]`` int16 / int32 / float32 as ``wavfile.read`` gives it).
    ``rir_loader`` may return ``None`` or raise ``ValueError`` for an unreadable file: the reference's zero RIR is used.
    """

    def __init__(self, device=None, n_samples: int = 16000, rir_max_taps: int = 16000, clip_capacity: int = 256,
                 rir_capacity: int = 4096, stage_bytes: int = 64 << 20, renderer: Optional[AudioRenderer] = None):
        self.rend = renderer or AudioRenderer(device)
        self.device = self.rend.device
        self.lib = self.rend.lib
        self.N = int(n_samples)
        self.Lmax = int(rir_max_taps) + (int(rir_max_taps) & 1)             # even rows: the packed-load variant of the kernel
        self.clips = RowCache(clip_capacity)
        self.rirs = RowCache(rir_capacity)
        dev = self.device
        self.clip_pool = torch.zeros((clip_capacity, self.N), dtype=torch.int16, device=dev)
        self.rir_pool = torch.zeros((rir_capacity, 2, self.Lmax), dtype=torch.float32, device=dev)
        self.rir_len = torch.zeros((rir_capacity,), dtype=torch.int32, device=dev)
        self._rir_len_host = np.zeros((rir_capacity,), dtype=np.int32)
        self._stage_host = torch.empty(stage_bytes, dtype=torch.uint8).pin_memory()
        self._stage_dev = torch.empty(stage_bytes, dtype=torch.uint8, device=dev)
        self._stage_free = torch.cuda.Event()
        self._stage_busy = False

    # ---- miss handling ------------------------------------------------------------------------------
    def _frames_of(self, src, what: str):
        """-> (uint8 view of the interleaved frames, format, channels, frames)."""
        if isinstance(src, np.ndarray):
            a = np.ascontiguousarray(src)
            if a.dtype not in _NP_FORMAT:
                raise _lib.M2HError(f"{what}: unsupported sample dtype {a.dtype} (wavfile.read gives int16 / int32 / float32)")
            ch = 1 if a.ndim == 1 else a.shape[1]
            return a.view(np.uint8).reshape(-1), _NP_FORMAT[a.dtype], ch, a.shape[0]
        info = parse_wav(src)
        buf = np.frombuffer(src, dtype=np.uint8)
        n = int(info.frames) * info.channels * _FORMAT_BYTES[info.format]
        return buf[info.data_offset: info.data_offset + n], info.format, info.channels, int(info.frames)

    def _flush_stage(self, jobs, used: int):
        """One H2D copy of the staged frames, then one ingest kernel per miss (all on the current stream)."""
        if not jobs:
            return
        st = _stream_ptr(self.device)
        self._stage_dev[:used].copy_(self._stage_host[:used], non_blocking=True)
        base = self._stage_dev.data_ptr()
        with torch.cuda.device(self.device):
            for kind, off, fmt, ch, frames, row in jobs:
                if kind == "rir":
                    rc = self.lib.m2h_rir_ingest(base + off, fmt, ch, frames, self.rir_pool[row].data_ptr(), self.Lmax, st)
                    _lib.check(rc, "m2h_rir_ingest")
                else:
                    rc = self.lib.m2h_clip_ingest(base + off, fmt, ch, frames, self.clip_pool[row].data_ptr(), self.N, st)
                    _lib.check(rc, "m2h_clip_ingest")
        self._stage_free.record(torch.cuda.current_stream(self.device))
        self._stage_busy = True

    def _stage(self, jobs, used, kind, frames_u8, fmt, ch, frames, row):
        n = frames_u8.size
        if n > self._stage_host.numel():
            raise _lib.M2HError(f"wav body of {n} bytes exceeds the staging buffer")
        if used + n > self._stage_host.numel():
            self._flush_stage(jobs, used)
            jobs.clear()
            used = 0
        if used == 0 and self._stage_busy:          # the pinned buffer is still the source of an earlier copy
            self._stage_free.synchronize()
            self._stage_busy = False
        if n:
            self._stage_host.numpy()[used: used + n] = frames_u8
        jobs.append((kind, used, fmt, ch, frames, row))
        return used + ((n + 15) & ~15)

    # ---- public API ---------------------------------------------------------------------------------------
    def render_keys(self, clip_keys: Sequence[Sequence[Hashable]], rir_keys: Sequence[Sequence[Hashable]],
                    clip_loader: Callable[[Hashable], Union[bytes, np.ndarray]],
                    rir_loader: Callable[[Hashable], RirSource], n_gt: int = 0, sim_dtypes: bool = False,
                    want_phase: bool = False, raw_mag: bool = False):
        """``[B][K]`` keys -> the dict :meth:`AudioRenderer.render` returns (mix_mag, mix_phase, gt, waves)."""
        B = len(clip_keys)
        K = len(clip_keys[0]) if B else 0
        if B == 0 or K == 0 or len(rir_keys) != B or any(len(r) != K for r in rir_keys) or any(len(c) != K for c in clip_keys):
            raise _lib.M2HError("render_keys: clip_keys and rir_keys must be non-empty [B][K] lists of equal shape")
        ci = np.empty((B, K), dtype=np.int32)
        ri = np.empty((B, K), dtype=np.int32)
        # Misses are transactional: a key is mapped to its row before the row is filled, so if anything below raises
        # (a loader's FileNotFoundError / OSError, a sample-rate assertion, a non-PCM16 clip, a cache smaller than the
        # batch) every key inserted by this call is unmapped again - a caller that catches the error and carries on
        # must never get a cache hit on a row that was not written.
        new_c, new_r = [], []
        try:
            return self._render_keys(clip_keys, rir_keys, clip_loader, rir_loader, n_gt, sim_dtypes, want_phase, raw_mag,
                                     B, K, ci, ri, new_c, new_r)
        except BaseException:
            for key in new_c:
                self.clips.forget(key)
            for key in new_r:
                self.rirs.forget(key)
            raise

    def _grow_rir_pool(self, frames: int):
        """RIRs longer than the pool rows (the reference accepts any length): re-allocate the pool with longer rows."""
        new_L = ((frames + 3999) // 4000) * 4000
        pool = torch.zeros((self.rir_pool.size(0), 2, new_L), dtype=torch.float32, device=self.device)
        pool[:, :, :self.Lmax].copy_(self.rir_pool)
        self.rir_pool, self.Lmax = pool, new_L

    def _render_keys(self, clip_keys, rir_keys, clip_loader, rir_loader, n_gt, sim_dtypes, want_phase, raw_mag,
                     B, K, ci, ri, new_c, new_r):
        jobs, used = [], 0
        pinned_c, pinned_r = set(), set()
        len_dirty = False
        for b in range(B):
            for k in range(K):
                row, hit = self.clips.lookup(clip_keys[b][k], pinned_c)
                if not hit:
                    new_c.append(clip_keys[b][k])
                    u8, fmt, ch, frames = self._frames_of(clip_loader(clip_keys[b][k]), "clip")
                    if fmt != WAV_PCM16:
                        raise _lib.M2HError("clips must be 16-bit PCM (the reference's 1 s / 16 kHz int16 sounds)")
                    used = self._stage(jobs, used, "clip", u8, fmt, ch, min(frames, self.N), row)
                ci[b, k] = row
                row, hit = self.rirs.lookup(rir_keys[b][k], pinned_r)
                if not hit:
                    new_r.append(rir_keys[b][k])
                    try:
                        src = rir_loader(rir_keys[b][k])
                    except ValueError:
                        logging.warning("{} file is not readable".format(rir_keys[b][k]))
                        src = None
                    frames = 0
                    if src is not None:
                        try:
                            u8, fmt, ch, frames = self._frames_of(src, "RIR")
                        except ValueError:
                            logging.warning("{} file is not readable".format(rir_keys[b][k]))
                            frames = 0
                    if frames > self.Lmax:
                        self._flush_stage(jobs, used)       # staged rows were addressed in the old pool
                        jobs, used = [], 0
                        self._grow_rir_pool(frames)
                    if frames > 0:                # frames == 0: unreadable / empty -> zero RIR = length 0 (simulator_train.py:411-415)
                        used = self._stage(jobs, used, "rir", u8, fmt, ch, frames, row)
                    self._rir_len_host[row] = frames
                    len_dirty = True
                ri[b, k] = row
        self._flush_stage(jobs, used)
        if len_dirty:
            self.rir_len.copy_(torch.from_numpy(self._rir_len_host), non_blocking=False)
        dev = self.device
        ci_d = torch.from_numpy(ci).to(dev, non_blocking=False)
        ri_d = torch.from_numpy(ri).to(dev, non_blocking=False)
        waves = torch.empty((B, K, 2, self.N), dtype=torch.int16, device=dev)
        with torch.cuda.device(dev):
            rc = self.lib.m2h_rir_convolve_indexed(self.clip_pool.data_ptr(), ci_d.data_ptr(), self.rir_pool.data_ptr(),
                                                   self.rir_len.data_ptr(), ri_d.data_ptr(), B, K, self.N, self.Lmax,
                                                   waves.data_ptr(), self.rend.workspace.data_ptr(),
                                                   self.rend.workspace.numel(), _stream_ptr(dev))
        _lib.check(rc, "m2h_rir_convolve_indexed")
        mix, phase, gt = self.rend.stft_waves(waves, n_gt, sim_dtypes, want_phase, raw_mag)
        return {"mix_mag": mix, "mix_phase": phase, "gt": gt, "waves": waves}

    def stats(self):
        return {"clip_hits": self.clips.hits, "clip_misses": self.clips.misses,
                "rir_hits": self.rirs.hits, "rir_misses": self.rirs.misses}
