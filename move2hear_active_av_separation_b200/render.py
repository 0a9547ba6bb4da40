"""Host side of the mixture render: batched device API + drop-in shims for the reference's call sites.

Mirrors (same names, argument meaning, return values, side effects, error behaviour):

* ``HabitatSimAudioEnabledTrain.get_current_mixed_bin_audio_mag_spec``
  (habitat_audio/simulator_train.py:386-486; eval twin simulator_eval.py:396-500), and
* ``PassiveDataset.compute_audiospects`` (audio_separation/pretrain/datasets/dataset.py:162-228).

All arithmetic runs in ``libm2h_b200.so`` (csrc/render_conv.cuh, csrc/render_stft.cuh) through the C ABI of
``include/m2h.h``; torch is used only for device memory and streams.  There is no CPU path.
"""
from __future__ import annotations

import logging
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

N_FFT = 1023
HOP_LENGTH = 512
N_BINS = 512
SAMPLE_RATE = 16000
FLAG_F16_OBS = 1
FLAG_RAW_MAG = 2
FLAG_GT_ONLY = 4
_NP_DTYPE = {torch.float16: np.float16, torch.float32: np.float32, torch.int16: np.int16, torch.int32: np.int32}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, name: str, dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.M2HError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise _lib.M2HError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


class AudioRenderer:
    """Owns the per-device workspace and launches the render kernels on the current torch stream."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.M2HError("no CUDA device: the B200 render path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.m2h_init(self.device.index), "m2h_init")
            nbytes = self.lib.m2h_render_workspace_bytes(self.device.index)
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    # ---- device-level API -------------------------------------------------------------------
    def rir_convolve(self, mono: torch.Tensor, rir: torch.Tensor, rir_len: torch.Tensor,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int16 ``[B,K,N]`` (*) fp32 ``[B,K,2,Lmax]`` -> rounded int16 waves ``[B,K,2,N]``."""
        mono = _need_cuda(mono, "mono", torch.int16)
        rir = _need_cuda(rir, "rir", torch.float32)
        rir_len = _need_cuda(rir_len, "rir_len", torch.int32)
        B, K, N = mono.shape
        if rir.shape[:3] != (B, K, 2) or rir_len.shape != (B, K):
            raise _lib.M2HError(f"shape mismatch: mono {tuple(mono.shape)} rir {tuple(rir.shape)} rir_len {tuple(rir_len.shape)}")
        Lmax = rir.shape[3]
        waves = out if out is not None else torch.empty((B, K, 2, N), dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.m2h_rir_convolve(mono.data_ptr(), rir.data_ptr(), rir_len.data_ptr(), B, K, N, Lmax,
                                           waves.data_ptr(), self.workspace.data_ptr(), self.workspace.numel(),
                                           _stream_ptr(self.device))
        _lib.check(rc, "m2h_rir_convolve")
        return waves

    def stft_waves(self, waves: torch.Tensor, n_gt: int = 0, sim_dtypes: bool = False, want_phase: bool = False,
                   raw_mag: bool = False, gt_only: bool = False):
        """STFT of the mixture (and of the first ``n_gt`` sources).  Returns ``(mix_mag, mix_phase|None, gt|None)``.
        ``gt_only``: only the per-source spectrograms (``mix_mag`` is None) - the mixture comes from another launch."""
        waves = _need_cuda(waves, "waves", torch.int16)
        B, K, _, N = waves.shape
        T = 1 + (N - 1) // HOP_LENGTH
        dt = torch.float16 if sim_dtypes else torch.float32
        mix = None if gt_only else torch.empty((B, N_BINS, T, 2), dtype=dt, device=self.device)
        phase = torch.empty((B, N_BINS, T, 2), dtype=dt, device=self.device) if want_phase and not gt_only else None
        gt = None
        if n_gt > 0:
            gt = torch.empty((B, n_gt, N_BINS, T, 4 if want_phase else 2), dtype=dt, device=self.device)
        flags = (FLAG_F16_OBS if sim_dtypes else 0) | (FLAG_RAW_MAG if raw_mag else 0) | (FLAG_GT_ONLY if gt_only else 0)
        with torch.cuda.device(self.device):
            rc = self.lib.m2h_stft_waves(waves.data_ptr(), B, K, N, n_gt, 1 if want_phase else 0, flags,
                                         mix.data_ptr() if mix is not None else None, phase.data_ptr() if phase is not None else None,
                                         gt.data_ptr() if gt is not None else None, _stream_ptr(self.device))
        _lib.check(rc, "m2h_stft_waves")
        return mix, phase, gt

    def stft_waves_x0(self, waves: torch.Tensor, out: Optional[torch.Tensor] = None, want_mag: bool = False):
        """STFT of the mixture written straight into the binaural separator's first-layer layout (``m2h_stft_waves_x0``) ->
        ``separator_cnn.PackedMixture``.  ``out``: a ZERO-INITIALISED bf16 ``[B,18,18,128]`` buffer to reuse (only interior
        cells are written; a fresh zeroed one is allocated otherwise).  ``want_mag``: also write the fp32 observation."""
        from .separator_cnn import PackedMixture
        waves = _need_cuda(waves, "waves", torch.int16)
        B, K, _, N = waves.shape
        if out is None:
            out = torch.zeros((B, 18, 18, 128), dtype=torch.bfloat16, device=self.device)
        pm = PackedMixture(out, torch.empty((B, N_BINS, 32, 2), dtype=torch.float32, device=self.device) if want_mag else None)
        if out.shape[0] != B or out.device != waves.device:
            raise _lib.M2HError(f"stft_waves_x0: out is {tuple(out.shape)} on {out.device} for a batch of {B} on {waves.device}")
        with torch.cuda.device(self.device):
            rc = self.lib.m2h_stft_waves_x0(waves.data_ptr(), B, K, N, out.data_ptr(), pm.mag.data_ptr() if want_mag else None,
                                            _stream_ptr(self.device))
        _lib.check(rc, "m2h_stft_waves_x0")
        return pm

    def render(self, mono, rir, rir_len, n_gt: int = 0, sim_dtypes: bool = False, want_phase: bool = False,
               raw_mag: bool = False):
        """conv + quantise + mix + STFT.  Returns dict(mix_mag, mix_phase, gt, waves)."""
        waves = self.rir_convolve(mono, rir, rir_len)
        mix, phase, gt = self.stft_waves(waves, n_gt, sim_dtypes, want_phase, raw_mag)
        return {"mix_mag": mix, "mix_phase": phase, "gt": gt, "waves": waves}

    def stft_f32(self, wave: torch.Tensor, want_phase: bool = False):
        """fp32 ``[B,C,N]`` -> (``|STFT|`` ``[B,512,T,C]``, phase or None)."""
        wave = _need_cuda(wave, "wave", torch.float32)
        B, C, N = wave.shape
        T = 1 + (N - 1) // HOP_LENGTH
        mag = torch.empty((B, N_BINS, T, C), dtype=torch.float32, device=self.device)
        ph = torch.empty_like(mag) if want_phase else None
        with torch.cuda.device(self.device):
            rc = self.lib.m2h_stft_f32(wave.data_ptr(), B, C, N, mag.data_ptr(),
                                       ph.data_ptr() if ph is not None else None, _stream_ptr(self.device))
        _lib.check(rc, "m2h_stft_f32")
        return mag, ph

    def istft(self, mag: torch.Tensor, phase: torch.Tensor, length: int = 16000) -> torch.Tensor:
        """fp32 magnitude / phase ``[B,512,T,C]`` (C = 1 or 2) -> waveforms ``[B,C,length]``: ``librosa.istft(mag * exp(1j *
        phase), hop_length=512, length=length)`` per channel, i.e. with librosa's inferred n_fft = 1022."""
        mag = _need_cuda(mag, "mag", torch.float32)
        phase = _need_cuda(phase, "phase", torch.float32)
        if mag.dim() != 4 or mag.shape != phase.shape or mag.shape[1] != N_BINS or mag.shape[3] not in (1, 2):
            raise _lib.M2HError(f"istft: expected mag / phase [B,512,T,1|2], got {tuple(mag.shape)} / {tuple(phase.shape)}")
        B, _, T, C = mag.shape
        out = torch.empty((B, C, length), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.m2h_istft(mag.data_ptr(), phase.data_ptr(), B, C, T, length, out.data_ptr(), _stream_ptr(self.device))
        _lib.check(rc, "m2h_istft")
        return out

    # ---- host-array convenience (what the shims use) -------------------------------------------------
    # The drop-in path runs once per environment step inside every env worker (14 processes per GPU, SURVEY hard part 6):
    # inputs are assembled in a pinned staging buffer and uploaded with ONE asynchronous copy, results come back through
    # ONE pinned buffer and ONE synchronisation (first version: three pageable uploads and four blocking .cpu() calls).
    def _pinned(self, name: str, nbytes: int) -> torch.Tensor:
        bufs = self.__dict__.setdefault("_pinned_bufs", {})
        buf = bufs.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = bufs[name] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
        return buf

    def pack_sources(self, monos: Sequence[np.ndarray], rirs: Sequence[np.ndarray]):
        """K int16 ``[N]`` clips + K wav-layout ``[L_k, 2]`` RIRs -> device tensors for a batch of one."""
        K = len(monos)
        N = len(monos[0])
        Lmax = max(1, max(len(r) for r in rirs))
        Lmax += Lmax & 1                                  # even rows: the packed-load variant of the kernel
        for m in monos:
            if np.asarray(m).dtype != np.int16:
                # the reference convolves whatever wavfile.read returned; its clips are int16 PCM - a silent cast of a
                # float clip would truncate it before the convolution
                raise _lib.M2HError(f"source clips must be int16 PCM (wavfile.read of the reference's sounds), got {np.asarray(m).dtype}")
        ev = self.__dict__.pop("_in_flight", None)
        if ev is not None:
            ev.synchronize()                              # the previous upload must have left the staging buffer (it long has)
        n_mono, n_rir, n_len = 2 * K * N, 4 * K * 2 * Lmax, 4 * K
        o_rir = (n_mono + 15) & ~15
        o_len = o_rir + n_rir
        total = o_len + n_len
        host = self._pinned("in", total)
        hb = host.numpy()
        mono_h = hb[:n_mono].view(np.int16).reshape(1, K, N)
        rir_h = hb[o_rir:o_rir + n_rir].view(np.float32).reshape(1, K, 2, Lmax)
        len_h = hb[o_len:o_len + n_len].view(np.int32).reshape(1, K)
        for k, (m, r) in enumerate(zip(monos, rirs)):
            mono_h[0, k] = m
            r = np.asarray(r)
            len_h[0, k] = len(r)
            if len(r):
                rir_h[0, k, :, :len(r)] = np.asarray(r, dtype=np.float32).T
            rir_h[0, k, :, len(r):] = 0
        dev_buf = torch.empty(total, dtype=torch.uint8, device=self.device)
        dev_buf.copy_(host[:total], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._in_flight = ev
        mono_d = dev_buf[:n_mono].view(torch.int16).view(1, K, N)
        rir_d = dev_buf[o_rir:o_rir + n_rir].view(torch.float32).view(1, K, 2, Lmax)
        len_d = dev_buf[o_len:o_len + n_len].view(torch.int32).view(1, K)
        return (mono_d, rir_d, len_d), N

    def fetch(self, tensors: Sequence[Optional[torch.Tensor]]):
        """Device tensors -> numpy arrays (``None`` passes through) through one pinned buffer and one synchronisation."""
        live = [t for t in tensors if t is not None]
        sizes = [t.numel() * t.element_size() for t in live]
        offs, total = [], 0
        for n in sizes:
            offs.append(total)
            total += (n + 15) & ~15
        host = self._pinned("out", total)
        for t, o, n in zip(live, offs, sizes):
            host[o:o + n].copy_(t.contiguous().view(-1).view(torch.uint8), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        out, it = [], iter(zip(live, offs, sizes))
        for t in tensors:
            if t is None:
                out.append(None)
                continue
            tt, o, n = next(it)
            a = host[o:o + n].numpy().view(_NP_DTYPE[tt.dtype]).reshape(tuple(tt.shape)).copy()    # fresh array owned by the caller
            out.append(a)
        return out

    def mono_gt_spec(self, mono: np.ndarray, gt_mono_mag_norm: float, want_phase: bool, dataset_semantics: bool = False):
        """Cached-per-sound monaural GT spectrogram (simulator_train.py:443-459, dataset.py:201-212).

        The simulator skips the RMS normalisation when ``GT_MONO_MAG_NORM == 0`` (``:452``); the dataset multiplies by it
        unconditionally (``dataset.py:208-209``: an all-zero target for the config default 0) - ``dataset_semantics``."""
        wave = torch.from_numpy(np.asarray(mono).astype("float32") / 32768).to(self.device)[None, None]
        mag, ph = self.stft_f32(wave, want_phase)
        mag = mag[0, :, :, 0]
        rms = torch.sqrt(torch.mean(mag * mag))
        if dataset_semantics or gt_mono_mag_norm != 0.0:
            # `if rms != 0` of the reference without a host round trip: an all-zero clip keeps its all-zero magnitude
            mag = torch.where(rms != 0, mag * gt_mono_mag_norm / rms, mag)
        mag, ph = self.fetch([mag, ph[0, :, :, 0] if want_phase else None])
        return mag, ph


_renderers = {}


def get_renderer(device=None) -> AudioRenderer:
    """Process-wide renderer per device, created lazily (safe with the reference's forkserver workers)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    key = (os.getpid(), idx)
    if key not in _renderers:
        _renderers[key] = AudioRenderer(torch.device("cuda", idx))
    return _renderers[key]


# ================================================================================================
# Drop-in replacements for the reference's methods (bind with `Class.method = function`)
# ================================================================================================

def istft(mag_l, phase_l, mag_r=None, phase_r=None, rend=None):
    r"""Drop-in for ``eval_metrics.istft`` (audio_separation/common/eval_metrics.py:232-252): ``[512, T]`` magnitude / phase
    arrays of a monaural or a binaural spectrogram -> list of one or two ``float32 [16000]`` waveforms.  Both ears are
    inverted in one pass on the GPU."""
    rend = rend or get_renderer()
    chans = [(mag_l, phase_l)]
    if mag_r is not None:
        assert phase_r is not None
        chans.append((mag_r, phase_r))
    mag = torch.from_numpy(np.stack([np.asarray(m, dtype=np.float32) for m, _ in chans], axis=-1)[None]).to(rend.device)
    ph = torch.from_numpy(np.stack([np.asarray(p, dtype=np.float32) for _, p in chans], axis=-1)[None]).to(rend.device)
    out = rend.istft(mag, ph, length=16000)[0].cpu().numpy()
    return [out[c] for c in range(len(chans))]


def _read_rir(path, sampling_rate):
    """wavfile.read with the reference's error convention (simulator_train.py:407-415, dataset.py:170-176)."""
    from scipy.io import wavfile
    try:
        sr, rir = wavfile.read(path)
        assert sr == sampling_rate, "RIR doesn't have sampling frequency of RIR_SAMPLING_RATE kHz"
    except ValueError:
        logging.warning("{} file is not readable".format(path))
        rir = np.zeros((sampling_rate, 2)).astype(np.float32)
    if len(rir) == 0:
        logging.debug("Empty RIR file at {}".format(path))
        rir = np.zeros((sampling_rate, 2)).astype(np.float32)
    return rir


def get_current_mixed_bin_audio_mag_spec(self):
    r"""Drop-in for ``HabitatSimAudioEnabledTrain/Eval.get_current_mixed_bin_audio_mag_spec``.

    Same file reads, caches (``_cached_mono_waveforms``, ``_cached_mono_specs``), side effects
    (``_gt_bin_audio_components``, ``_gt_mono_audio_components``, ``_mixed_bin_audio_phase``) and return value
    (``np.float16 [512, 32, 2]`` = log1p of the fp16 magnitude); the arithmetic runs on the GPU.
    """
    from scipy.io import wavfile
    cfg = self.config.AUDIO
    rend = get_renderer()
    self._gt_bin_audio_components = []
    self._gt_mono_audio_components = []
    monos, rirs = [], []
    for indx, source_position_index in enumerate(self._source_position_indices):
        name = self._current_sound_names[indx]
        if name not in self._cached_mono_waveforms:
            _, wav = wavfile.read(os.path.join(cfg.MONO_DIR, name + ".wav"))
            self._cached_mono_waveforms[name] = wav
        monos.append(self._cached_mono_waveforms[name])
        rir_file = os.path.join(cfg.RIR_DIR, self.current_scene_name, str(self.azimuth_angle),
                                str(self._receiver_position_index) + "_" + str(source_position_index) + ".wav")
        rirs.append(_read_rir(rir_file, cfg.RIR_SAMPLING_RATE))
    return render_sim_arrays(self, monos, rirs, list(self._current_sound_names), cfg.GT_MONO_MAG_NORM, rend)


def get_current_mixed_bin_audio_mag_spec_cached(self):
    r"""Variant of :func:`get_current_mixed_bin_audio_mag_spec` that keeps the clips and RIRs in device-side LRU caches
    (``ingest.CachedAudioRenderer``): a RIR wav is read, decoded and uploaded only the first time its
    ``(scene, azimuth, receiver, source)`` is rendered.  Same return value and side effects."""
    from scipy.io import wavfile
    from .ingest import CachedAudioRenderer
    cfg = self.config.AUDIO
    cached = getattr(self, "_m2h_cached_renderer", None)
    if cached is None:
        cached = self._m2h_cached_renderer = CachedAudioRenderer(renderer=get_renderer(), rir_max_taps=cfg.RIR_SAMPLING_RATE)
    names = list(self._current_sound_names)

    def clip_loader(name):
        if name not in self._cached_mono_waveforms:
            _, wav = wavfile.read(os.path.join(cfg.MONO_DIR, name + ".wav"))
            self._cached_mono_waveforms[name] = wav
        return self._cached_mono_waveforms[name]

    def rir_loader(key):
        scene, azimuth, receiver, source = key
        path = os.path.join(cfg.RIR_DIR, scene, str(azimuth), str(receiver) + "_" + str(source) + ".wav")
        with open(path, "rb") as f:
            data = f.read()
        from .ingest import parse_wav
        info = parse_wav(data)                      # ValueError -> "file is not readable" -> zero RIR, like the reference
        assert info.sample_rate == cfg.RIR_SAMPLING_RATE, "RIR doesn't have sampling frequency of RIR_SAMPLING_RATE kHz"
        return data

    rir_keys = [(self.current_scene_name, self.azimuth_angle, self._receiver_position_index, src)
                for src in self._source_position_indices]
    K = len(names)
    out = cached.render_keys([names], [rir_keys], clip_loader, rir_loader, n_gt=K, sim_dtypes=True, want_phase=True)
    monos = [clip_loader(n) for n in names]
    return _finish_sim_render(self, out, monos, names, cfg.GT_MONO_MAG_NORM, cached.rend)


def render_sim_arrays(self, monos, rirs, names, gt_mono_mag_norm, rend=None):
    """Array-level body of the simulator shim (also what the tests drive): sets the side-effect fields on ``self``."""
    rend = rend or get_renderer()
    K = len(monos)
    (mono_d, rir_d, len_d), _ = rend.pack_sources(monos, rirs)
    out = rend.render(mono_d, rir_d, len_d, n_gt=K, sim_dtypes=True, want_phase=True)
    return _finish_sim_render(self, out, monos, names, gt_mono_mag_norm, rend)


def _finish_sim_render(self, out, monos, names, gt_mono_mag_norm, rend):
    """Side-effect fields + return value of the simulator render from the device results of one mixture."""
    K = len(monos)
    gt, phase, mix = rend.fetch([out["gt"][0], out["mix_phase"][0], out["mix_mag"][0]])      # one pinned D2H + one sync
    self._gt_bin_audio_components = [gt[k] for k in range(K)]                                   # [512,T,4] fp16 each
    self._gt_mono_audio_components = []
    for k, name in enumerate(names):
        if name not in self._cached_mono_specs:
            if gt_mono_mag_norm != 0.0:
                assert getattr(getattr(self, "config", None), "AUDIO", None) is None or self.config.AUDIO.NORM_TYPE == "l2"
            mag, ph = rend.mono_gt_spec(monos[k], gt_mono_mag_norm, want_phase=True)
            self._cached_mono_specs[name] = np.stack([mag, ph], axis=-1)
        self._gt_mono_audio_components.append(self._cached_mono_specs[name].astype(np.float16))
    self._mixed_bin_audio_phase = phase
    return mix


def compute_audiospects(self, rirs_audio, target_class):
    r"""Drop-in for ``PassiveDataset.compute_audiospects`` (dataset.py:162-228).

    Returns ``(log1p(mixed_mag) f32 [512,32,2], [gt_bin_mag f32 [512,32,2]], [gt_mono_mag f32 [512,32,1]],
    np.int64 [1])``; keeps ``self._gt_mono_mag_cache`` and the zero-RIR fallback.
    """
    monos, rirs, names = [], [], []
    for rir_audio in rirs_audio:
        rirs.append(_read_rir(os.path.join(self.binaural_rir_dir, rir_audio[0]), self.rir_sampling_rate))
        monos.append(self.file2audio_dict[rir_audio[1]])
        names.append(rir_audio[1])
    return compute_audiospects_arrays(self, monos, rirs, names, target_class, self.audio_cfg.GT_MONO_MAG_NORM)


def compute_audiospects_arrays(self, monos, rirs, names, target_class, gt_mono_mag_norm, rend=None):
    rend = rend or get_renderer()
    (mono_d, rir_d, len_d), _ = rend.pack_sources(monos, rirs)
    out = rend.render(mono_d, rir_d, len_d, n_gt=1, sim_dtypes=False, want_phase=False)
    gt_bin, mix = rend.fetch([out["gt"][0, 0], out["mix_mag"][0]])
    gt_bin_mag = [gt_bin]
    if names[0] not in self._gt_mono_mag_cache:
        mag, _ = rend.mono_gt_spec(monos[0], gt_mono_mag_norm, want_phase=False, dataset_semantics=True)
        self._gt_mono_mag_cache[names[0]] = mag
    gt_mono_mag = [np.expand_dims(self._gt_mono_mag_cache[names[0]], axis=2).astype("float32")]
    return mix, gt_bin_mag, gt_mono_mag, np.array([target_class]).astype("int64")
