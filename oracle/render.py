"""CPU oracle: binaural mixture render (RIR convolution + quantise + mix + STFT).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  The reference has no
tests of its own; this restatement is pinned against OUTPUTS OF THE REFERENCE'S
OWN FUNCTION BODIES (``simulator_train.py:386-486`` + getters, the eval twin,
``dataset.py:162-228``), executed unmodified under ``sys.modules`` stubs on real
wav files by ``oracle/gen_ref_golden.py`` -> ``tests/golden/ref_render_*.npz``,
reproduced bit for bit by ``tests/test_oracle.py::test_oracle_matches_reference_functions``
(the only stand-in on that route is the librosa shim = :func:`stft_librosa`,
itself pinned against ``torch.stft`` in float64), and against float64
``np.convolve`` as an independent route for the convolution.

Restates, with in-memory arrays instead of ``wavfile.read``:

* ``habitat_audio/simulator_train.py:386-486``
  (``HabitatSimAudioEnabledTrain.get_current_mixed_bin_audio_mag_spec``; the
  eval twin ``habitat_audio/simulator_eval.py:396-500`` is the same math), and
* ``audio_separation/pretrain/datasets/dataset.py:162-228``
  (``PassiveDataset.compute_audiospects``).

Third-party arithmetic that is not vendored under /root/reference:

* ``scipy.signal.fftconvolve`` (reference pins scipy==1.5.4, requirements.txt:170):
  called directly - scipy 1.18.1 here, same pocketfft ``rfftn``/``irfftn`` +
  ``next_fast_len`` + centred crop.
* ``librosa.stft`` / ``librosa.magphase`` (librosa==0.8.0, requirements.txt:85):
  not installable here; restated from its published algorithm in
  :func:`stft_librosa` / :func:`magphase_librosa`.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve, get_window

N_FFT = 1023          # dataset.py:21, simulator_train.py:426 (hard-coded)
HOP_LENGTH = 512      # dataset.py:20
SAMPLE_RATE = 16000   # config/default.py:200 RIR_SAMPLING_RATE
N_BINS = 1 + N_FFT // 2   # 512


def n_frames(n_samples: int) -> int:
    """Frame count of librosa.stft(center=True) with the odd n_fft = 1023: 1 + (len(y) + 2*511 - 1023) // hop."""
    return 1 + (n_samples + 2 * (N_FFT // 2) - N_FFT) // HOP_LENGTH


def hann_periodic(n_fft: int = N_FFT) -> np.ndarray:
    """librosa.filters.get_window('hann', n_fft, fftbins=True) (float64)."""
    return get_window("hann", n_fft, fftbins=True)


def stft_librosa(y: np.ndarray, n_fft: int = N_FFT, hop_length: int = HOP_LENGTH) -> np.ndarray:
    """librosa-0.8.0 ``stft(y, n_fft, hop_length)`` with its defaults.

    win_length = n_fft, periodic Hann, center=True, pad_mode='reflect',
    dtype=complex64.  librosa multiplies the (float32) frames by the float64
    window, runs ``np.fft.rfft`` along the frame axis in float64, and stores
    the result into a complex64 array.  Returns ``[1 + n_fft//2, n_frames]``.
    """
    y = np.asarray(y)
    assert y.ndim == 1
    window = hann_periodic(n_fft).reshape(-1, 1)              # pad_center is a no-op (win_length == n_fft)
    yp = np.pad(y, int(n_fft // 2), mode="reflect")
    n_fr = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_fr)[None, :]
    frames = yp[idx]                                          # [n_fft, n_frames] (librosa.util.frame, axis=-1)
    out = np.empty((1 + n_fft // 2, n_fr), dtype=np.complex64, order="F")
    out[:] = np.fft.rfft(window * frames, axis=0)
    return out


def istft_librosa(stft_matrix: np.ndarray, hop_length: int = HOP_LENGTH, length: int = SAMPLE_RATE) -> np.ndarray:
    """librosa-0.8.0 ``istft(stft_matrix, hop_length=512, length=16000)`` as called by
    audio_separation/common/eval_metrics.py:232-252.

    librosa infers ``n_fft = 2 * (n_bins - 1)`` = **1022** from the 512 bins (not the 1023 the forward transform used),
    ``win_length = n_fft``, periodic Hann, ``center=True``: every frame is ``irfft(column, n=1022)`` (float64; the imaginary
    parts of bins 0 and 511 are ignored) times the window, overlap-added at ``512 * t``; the sum is divided by the window's
    sum of squares where that exceeds ``tiny``, the first ``n_fft // 2`` samples are dropped and the result is cut /
    zero-padded to ``length``; float32 output for a complex64 input.
    """
    stft_matrix = np.asarray(stft_matrix)
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    n_fr = stft_matrix.shape[1]
    window = get_window("hann", n_fft, fftbins=True)
    expected = n_fft + hop_length * (n_fr - 1)
    y = np.zeros(expected, dtype=np.float32)
    for t in range(n_fr):
        y[t * hop_length: t * hop_length + n_fft] += (window * np.fft.irfft(stft_matrix[:, t], n=n_fft)).astype(np.float32)
    wss = np.zeros(expected, dtype=np.float64)
    for t in range(n_fr):
        wss[t * hop_length: t * hop_length + n_fft] += window ** 2
    wss = wss.astype(np.float32)                       # librosa.filters.window_sumsquare(dtype = the output dtype)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    out = np.zeros(length, dtype=np.float32)
    out[: min(length, len(y))] = y[:length]
    return out


def istft_eval(mag_l, phase_l, mag_r=None, phase_r=None):
    """``eval_metrics.istft`` (audio_separation/common/eval_metrics.py:232-252): list of 1 or 2 float32 waveforms."""
    sig = [istft_librosa(np.asarray(mag_l) * np.exp(1j * np.asarray(phase_l)))]
    if mag_r is not None:
        assert phase_r is not None
        sig.append(istft_librosa(np.asarray(mag_r) * np.exp(1j * np.asarray(phase_r))))
    return sig


def magphase_librosa(D: np.ndarray):
    """librosa-0.8.0 ``magphase(D, power=1)``: ``(|D|, exp(1j * angle(D)))``."""
    mag = np.abs(D)
    phase = np.exp(1.0j * np.angle(D))
    return mag, phase


def log1p_f16(x16: np.ndarray) -> np.ndarray:
    """``np.log1p`` of a float16 array with the reference's pinned numpy semantics.

    numpy 1.19.5 (requirements.txt) evaluates half-precision ufuncs as
    ``float -> log1pf -> half``.  numpy >= 2 on AVX512-FP16 hosts dispatches to an
    SVML half kernel that is a few fp16 ulp off (41 % of inputs differ in this
    container), which would make the oracle depend on the host CPU - so the
    float32 route is spelled out.
    """
    assert x16.dtype == np.float16
    return np.log1p(x16.astype(np.float32)).astype(np.float16)


def convolve_quantise(mono: np.ndarray, rir: np.ndarray) -> np.ndarray:
    """simulator_train.py:417-423 / dataset.py:178-185.

    ``mono``: int16 ``[N]``; ``rir``: float32 ``[L, 2]`` (wav layout).  Returns
    float32 ``[2, N]``: fftconvolve(mode="same") per ear, np.round (half to
    even), astype int16 (wraps modulo 2**16), astype float32, times 1/32768.
    A zero-length RIR is replaced by 16000 zero taps (simulator_train.py:413-415).
    """
    if len(rir) == 0:
        rir = np.zeros((SAMPLE_RATE, 2), dtype=np.float32)
    conv = []
    for ch in range(rir.shape[-1]):
        conv.append(fftconvolve(mono, rir[:, ch], mode="same"))
    conv = np.array(conv)
    with np.errstate(invalid="ignore"):
        conv = np.round(conv).astype("int64").astype("int16").astype("float32")
    conv *= (1 / 32768)
    return conv


def mono_gt_mag(mono: np.ndarray, gt_mono_mag_norm: float = 1.2, with_phase: bool = False,
                dataset_semantics: bool = False):
    """GT monaural spectrogram of a source clip (cached per sound in the reference).

    simulator_train.py:443-459 / dataset.py:201-212: STFT of ``mono/32768``
    (float32), magnitude RMS-normalised to ``GT_MONO_MAG_NORM``.  The two call
    sites differ when the norm is 0 (the config default): the simulator skips
    the normalisation (``:452``), the dataset multiplies by it unconditionally
    (``dataset.py:208-209``: zeros) - found by running the reference's own code
    (``tests/golden/ref_render_k2_short_nonorm.npz``).
    """
    m = mono.astype("float32") / 32768
    D = stft_librosa(m)
    mag, ph = magphase_librosa(D)
    rms = np.power(np.mean(np.power(mag, 2)), 0.5)
    if (dataset_semantics or gt_mono_mag_norm != 0.0) and rms != 0.0:
        mag = mag * gt_mono_mag_norm / rms
    if with_phase:
        return mag, np.angle(ph)
    return mag


def render_sim(monos, rirs, gt_mono_mag_norm: float = 1.2):
    """simulator_train.py:386-486 with arrays in place of wav files.

    ``monos``: K int16 ``[N]`` arrays; ``rirs``: K float32 ``[L_k, 2]`` arrays.
    Returns a dict with the return value (``mixed_bin_audio_mag``: float16
    ``[512, T, 2]`` = log1p of the fp16 magnitude) and the three side-effect
    fields the sensors read (task.py:89-176).
    """
    gt_bin, gt_mono = [], []
    mixed = 0
    waves = []
    for mono, rir in zip(monos, rirs):
        conv = convolve_quantise(mono, rir)
        waves.append(conv)
        comps = []
        for ch in range(2):
            D = stft_librosa(np.asfortranarray(conv[ch]))
            mag, ph = magphase_librosa(D)
            comps += [mag, np.angle(ph)]
        gt_bin.append(np.stack(comps, axis=-1).astype(np.float16))          # [512,T,4] = mag_l, ph_l, mag_r, ph_r
        mmag, mph = mono_gt_mag(mono, gt_mono_mag_norm, with_phase=True)
        gt_mono.append(np.stack([mmag, mph], axis=-1).astype(np.float16))   # [512,T,2]
        mixed = mixed + conv
    mixed = mixed / np.float32(len(monos))
    mags, phases = [], []
    for ch in range(2):
        D = stft_librosa(np.asfortranarray(mixed[ch]))
        mag, ph = magphase_librosa(D)
        mags.append(mag)
        phases.append(np.angle(ph))
    mixed_mag32 = np.stack(mags, axis=-1)
    mixed_mag = mixed_mag32.astype(np.float16)
    return {
        "mixed_bin_audio_mag": log1p_f16(mixed_mag),                         # float16 [512,T,2]
        "mixed_bin_audio_phase": np.stack(phases, axis=-1).astype(np.float16),
        "gt_bin_audio_components": gt_bin,
        "gt_mono_audio_components": gt_mono,
        # extras for parity checks (not part of the reference's return values)
        "_mixed_wave": mixed.astype(np.float32),
        "_source_waves": np.stack(waves).astype(np.float32),
        "_mixed_mag_f32": mixed_mag32.astype(np.float32),
    }


def render_dataset(monos, rirs, target_class: int, gt_mono_mag_norm: float = 1.2):
    """dataset.py:162-228 (``compute_audiospects``) with arrays in place of files.

    Returns ``(log1p(mixed_mag) f32 [512,T,2], [gt_bin_mag f32 [512,T,2]],
    [gt_mono_mag f32 [512,T,1]], int64 [1])`` plus a dict of extras.
    """
    gt_mono_mag, gt_bin_mag = [], []
    mixed = 0
    for idx, (mono, rir) in enumerate(zip(monos, rirs)):
        conv = convolve_quantise(mono, rir)
        if idx == 0:
            mags = []
            for ch in range(2):
                mag, _ = magphase_librosa(stft_librosa(np.asfortranarray(conv[ch])))
                mags.append(mag)
            gt_bin_mag.append(np.stack(mags, axis=-1).astype("float32"))
            mm = mono_gt_mag(mono, gt_mono_mag_norm, dataset_semantics=True)
            gt_mono_mag.append(np.expand_dims(mm, axis=2).astype("float32"))
        mixed = mixed + conv
    mixed = mixed / np.float32(len(monos))
    mags = []
    for ch in range(2):
        mag, _ = magphase_librosa(stft_librosa(np.asfortranarray(mixed[ch])))
        mags.append(mag)
    mixed_mag = np.stack(mags, axis=-1).astype("float32")
    extras = {"_mixed_wave": mixed.astype(np.float32), "_mixed_mag_f32": mixed_mag}
    return (np.log1p(mixed_mag), gt_bin_mag, gt_mono_mag,
            np.array([target_class]).astype("int64"), extras)


# --------------------------------------------------------------------------
# Seeded synthetic inputs - shared by the tests,
# the golden-vector generator and bench.py so that every arm sees the same data.
# --------------------------------------------------------------------------

def synth_monos(B: int, K: int, N: int = SAMPLE_RATE, seed: int = 0) -> np.ndarray:
    """int16 ``[B,K,N]``: round(3000 * N(0,1)) clipped to +-32767."""
    rng = np.random.default_rng(seed)
    x = np.round(3000.0 * rng.standard_normal((B, K, N)))
    return np.clip(x, -32767, 32767).astype(np.int16)


def synth_rirs(B: int, K: int, L: int = SAMPLE_RATE, seed: int = 1, ragged: bool = False,
               Lmin: int = 4000, gain: float = 0.05):
    """float32 ``[B,K,2,L]`` (ear-planar) + int32 lengths ``[B,K]``.

    ``N(0,1) * exp(-t/tau) * gain`` with ``tau = L_k/6``; ragged draws
    ``L_k ~ U{Lmin..L}`` and zero-fills the tail.
    """
    rng = np.random.default_rng(seed)
    lens = np.full((B, K), L, dtype=np.int32)
    if ragged:
        lens = rng.integers(Lmin, L + 1, size=(B, K)).astype(np.int32)
    rir = rng.standard_normal((B, K, 2, L)).astype(np.float32)
    t = np.arange(L, dtype=np.float32)
    for b in range(B):
        for k in range(K):
            lk = int(lens[b, k])
            env = np.exp(-t / (lk / 6.0)).astype(np.float32) * np.float32(gain)
            rir[b, k] *= env
            rir[b, k, :, lk:] = 0
    return rir, lens


def rir_wav_layout(rir_planar: np.ndarray, length: int) -> np.ndarray:
    """``[2, Lmax]`` planar -> the reference's wav layout ``[L, 2]``."""
    return np.ascontiguousarray(rir_planar[:, :length].T)
