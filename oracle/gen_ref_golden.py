"""Pin ``oracle/render.py`` against the reference's OWN function bodies.

Run in the BUILD container only (needs /root/reference):

    python -m oracle.gen_ref_golden

The reference's render functions live in modules that import habitat / habitat_sim / librosa, none of which is
installable here.  They are imported UNMODIFIED, by path, under ``sys.modules`` stubs:

* ``habitat_audio/simulator_train.py`` and ``habitat_audio/simulator_eval.py`` - stubs for ``habitat.*``,
  ``habitat_sim.*`` (class / decorator plumbing only: no arithmetic lives there, SURVEY 8c iv) and ``librosa``;
  the unbound ``get_current_mixed_bin_audio_mag_spec`` (``simulator_train.py:386-486``,
  ``simulator_eval.py:396-500``) and the three getters (``:363-384``) are then called on a plain attribute bag that
  carries exactly the fields the function reads (config.AUDIO.*, current_scene_name, azimuth_angle, ...).
* ``audio_separation/pretrain/datasets/dataset.py`` - needs only the ``librosa`` stub; the unbound
  ``PassiveDataset.compute_audiospects`` (``:162-228``) is called the same way.

Inputs are REAL wav files written to a temp dir with ``scipy.io.wavfile.write`` (PCM16 clips, IEEE-float stereo RIRs,
one unreadable RIR, one zero-length RIR), so the reference's own ``wavfile.read`` / ``fftconvolve`` / ``np.round`` /
cast / cache / fallback lines all execute.  The only stand-in arithmetic is the ``librosa`` shim
(``stft`` / ``magphase`` = ``oracle.render.stft_librosa`` / ``magphase_librosa``, the librosa-0.8.0 restatement, which
is pinned separately against ``torch.stft`` in float64 by ``tests/test_oracle.py``).

One host dependence is recorded rather than hidden: the simulator returns ``np.log1p(float16 array)``; numpy >= 2 on
AVX512-FP16 hosts evaluates that with an SVML half kernel that differs by one fp16 ulp from the reference's pinned
numpy 1.19.5 (float -> log1pf -> half) on a large share of inputs.  The generator therefore records both the
reference's return value as computed on this host (``sim_return_hostnumpy``) and the argument the reference passed
to ``np.log1p`` (``sim_log1p_arg_f16``, captured by wrapping ``np.log1p`` for the duration of the call).

Outputs: ``tests/golden/ref_render_*.npz``; ``tests/test_oracle.py::test_oracle_matches_reference_functions`` asserts
the oracle reproduces them bit for bit from the same seeded inputs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
from scipy.io import wavfile

from . import render as R
from .gen_golden import GOLDEN, render_case_inputs

REF_ROOT = "/root/reference"

# (name, K, RIR lengths (-1 = unreadable file, 0 = zero-length wav), seed, gt_mono_mag_norm)
REF_CASES = [
    ("k2_L16000", 2, [16000, 16000], 21, 1.2),
    ("k2_ragged", 2, [7999, 4000], 22, 1.2),
    ("k2_unreadable_empty", 2, [-1, 0], 23, 1.2),
    ("k2_short_nonorm", 2, [1, 2], 24, 0.0),
    ("k3_long", 3, [20000, 32000, 16001], 25, 1.2),
]


# ------------------------------------------------------------------ stubs

def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Registry:
    """habitat.core.registry.registry: only the decorators the two simulator modules use."""

    def register_simulator(self, *a, **k):
        return lambda cls: cls


def install_stubs():
    """Put the stub modules into ``sys.modules`` (idempotent).  No arithmetic except the librosa shim."""
    _module("librosa", stft=lambda y, hop_length, n_fft: R.stft_librosa(y, n_fft=n_fft, hop_length=hop_length),
            magphase=R.magphase_librosa,
            istft=lambda D, hop_length, length: R.istft_librosa(D, hop_length=hop_length, length=length))
    _module("habitat")
    _module("habitat.core")
    _module("habitat.core.registry", registry=_Registry())
    _module("habitat.core.simulator", Config=object, AgentState=object, ShortestPathPoint=object)
    _module("habitat.sims")
    _module("habitat.sims.habitat_simulator")
    _module("habitat.sims.habitat_simulator.habitat_simulator", HabitatSim=type("HabitatSim", (), {}))
    _module("habitat.sims.habitat_simulator.actions", HabitatSimActions=object)
    _module("habitat_sim", AgentState=object)
    _module("habitat_sim.utils")
    _module("habitat_sim.utils.common", quat_from_angle_axis=None, quat_from_coeffs=None, quat_to_angle_axis=None)
    # the real package __init__ imports every habitat-dependent module: register an empty parent instead
    pkg = _module("habitat_audio")
    pkg.__path__ = []
    _load("habitat_audio.utils", "habitat_audio/utils.py")       # torch/numpy only: the real file


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    install_stubs()
    sim_train = _load("habitat_audio.simulator_train", "habitat_audio/simulator_train.py")
    sim_eval = _load("habitat_audio.simulator_eval", "habitat_audio/simulator_eval.py")
    dataset = _load("ref_passive_dataset", "audio_separation/pretrain/datasets/dataset.py")
    return sim_train, sim_eval, dataset


# ------------------------------------------------------------------ wav fixtures

def write_case_files(tmp, K, lens, seed):
    """Seeded clips + RIRs as the wav files the reference reads.  Returns (monos, rirs as the oracle takes them)."""
    eff = [max(l, 0) for l in lens]
    monos, rirs, _, _ = render_case_inputs(K, eff, seed)
    mono_dir = os.path.join(tmp, "mono")
    rir_dir = os.path.join(tmp, "rir", "scene0", "90")
    os.makedirs(mono_dir, exist_ok=True)
    os.makedirs(rir_dir, exist_ok=True)
    for k in range(K):
        wavfile.write(os.path.join(mono_dir, "sound%d.wav" % k), R.SAMPLE_RATE, monos[k])
        path = os.path.join(rir_dir, "7_%d.wav" % (100 + k))
        if lens[k] < 0:
            with open(path, "wb") as f:            # not a RIFF file: wavfile.read raises ValueError
                f.write(b"JUNKJUNKJUNKJUNKJUNKJUNKJUNKJUNKJUNKJUNKJUNKJUNK")
            rirs[k] = np.zeros((0, 2), np.float32)
        else:
            wavfile.write(path, R.SAMPLE_RATE, np.ascontiguousarray(rirs[k], dtype=np.float32).reshape(-1, 2))
    return monos, rirs


def fake_sim(tmp, K, norm):
    """Exactly the attributes ``get_current_mixed_bin_audio_mag_spec`` reads from ``self``."""
    audio = SimpleNamespace(MONO_DIR=os.path.join(tmp, "mono"), RIR_DIR=os.path.join(tmp, "rir"),
                            RIR_SAMPLING_RATE=R.SAMPLE_RATE, GT_MONO_MAG_NORM=norm, NORM_TYPE="l2")
    return SimpleNamespace(config=SimpleNamespace(AUDIO=audio), current_scene_name="scene0", azimuth_angle=90,
                           _receiver_position_index=7, _source_position_indices=[100 + k for k in range(K)],
                           _current_sound_names=["sound%d" % k for k in range(K)],
                           _cached_mono_waveforms={}, _cached_mono_specs={},
                           _gt_bin_audio_components=[], _gt_mono_audio_components=[], _mixed_bin_audio_phase=None)


class _CaptureLog1p:
    """Record the argument of ``np.log1p`` while the reference function runs (the ufunc itself still computes)."""

    def __enter__(self):
        self.args = []
        self._orig = np.log1p

        def wrapped(x, *a, **k):
            self.args.append(np.array(x, copy=True))
            return self._orig(x, *a, **k)
        np.log1p = wrapped
        return self

    def __exit__(self, *exc):
        np.log1p = self._orig


def run_sim(cls, tmp, K, norm, calls=2):
    """Call the reference method ``calls`` times on one attribute bag (2nd call = clip / mono-spec cache hits)."""
    sim = fake_sim(tmp, K, norm)
    outs = []
    for _ in range(calls):
        with _CaptureLog1p() as cap:
            ret = cls.get_current_mixed_bin_audio_mag_spec(sim)
        outs.append({
            "sim_return_hostnumpy": ret,
            "sim_log1p_arg_f16": cap.args[-1],
            "sim_mixed_phase_f16": cls.get_current_mixed_bin_audio_phase_spec(sim),
            "sim_gt_bin_f16": cls.get_current_gt_bin_audio_components(sim),       # [512,T,4K]
            "sim_gt_mono_f16": cls.get_current_gt_mono_audio_components(sim),     # [512,T,2K]
        })
    return outs, sim


def run_dataset(ds_mod, tmp, K, norm, monos, target_class=3, calls=2):
    ds = SimpleNamespace(binaural_rir_dir=os.path.join(tmp, "rir"), rir_sampling_rate=R.SAMPLE_RATE,
                         audio_cfg=SimpleNamespace(GT_MONO_MAG_NORM=norm), _gt_mono_mag_cache={},
                         file2audio_dict={"sound%d.wav" % k: monos[k] for k in range(K)})
    rirs_audio = [(os.path.join("scene0", "90", "7_%d.wav" % (100 + k)), "sound%d.wav" % k) for k in range(K)]
    outs = []
    for _ in range(calls):
        with _CaptureLog1p() as cap:
            logmag, gt_bin, gt_mono, tc = ds_mod.PassiveDataset.compute_audiospects(ds, rirs_audio, target_class)
        outs.append({"ds_return_hostnumpy": logmag, "ds_log1p_arg_f32": cap.args[-1], "ds_gt_bin_mag_f32": gt_bin[0],
                     "ds_gt_mono_mag_f32": gt_mono[0], "ds_target_class": tc})
    return outs


def postproc_case_inputs(B=3, seed=71, T=32):
    """Seeded inputs of the mask consumers (STFT_L2_distance, AcousticMem, compute_waveform_quality)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    d = {"mixed_audio": torch.log1p(torch.rand(B, 512, T, 2, generator=g) * 3.0),
         "masks": torch.randn(B, 512, T, 2, generator=g) * 0.5 + 0.3,
         "gt_bin_comps": torch.stack([torch.rand(B, 512, T, generator=g) * 2, (torch.rand(B, 512, T, generator=g) * 2 - 1) * 3.14159,
                                      torch.rand(B, 512, T, generator=g) * 2, (torch.rand(B, 512, T, generator=g) * 2 - 1) * 3.14159], dim=-1),
         "pred_mono": torch.rand(B, 512, T, 1, generator=g) * 2,
         "gt_mono_comps": torch.stack([torch.rand(B, 512, T, generator=g) * 2, (torch.rand(B, 512, T, generator=g) * 2 - 1) * 3.14159], dim=-1),
         "prev_mem": torch.rand(B, 512, T, 1, generator=g) * 2}
    return d


def waveform_case_inputs(seed=81):
    """A consistent spectrogram set for compute_waveform_quality: STFTs of seeded waveforms (simulator dtypes widened)."""
    rng = np.random.default_rng(seed)
    gt = (rng.standard_normal(16000) * 0.1).astype(np.float32)
    other = (rng.standard_normal((2, 16000)) * 0.1).astype(np.float32)
    mixed = np.stack([0.6 * gt + other[0], 0.5 * gt + other[1]]).astype(np.float32)
    def spec(y):
        D = R.stft_librosa(y)
        return np.abs(D).astype(np.float32), np.angle(D).astype(np.float32)
    ml, pl = spec(mixed[0]); mr, pr = spec(mixed[1]); gm, gp = spec(gt)
    noise = np.abs(rng.standard_normal(gm.shape)).astype(np.float32)
    return {"mixed_bin_audio_mag": np.stack([ml, mr], -1)[None], "mixed_bin_audio_phase": np.stack([pl, pr], -1)[None],
            "gt_mono_mag": gm[None, :, :, None], "gt_mono_phase": gp[None, :, :, None],
            "pred_mono": (gm * 0.8 + 0.3 * noise)[None, :, :, None], "pred_monoFromMem": (gm * 0.9 + 0.1 * noise)[None, :, :, None]}


def gen_postproc():
    """tests/golden/postproc.npz: outputs of the reference's own STFT_L2_distance, AcousticMem (both variants) and
    compute_waveform_quality (eval_metrics.py / memory_nets.py imported unmodified)."""
    import torch
    install_stubs()
    em = _load("ref_eval_metrics", "audio_separation/common/eval_metrics.py")
    mn = _load("ref_memory_nets", "audio_separation/rl/models/memory_nets.py")
    d = postproc_case_inputs()
    out = {}
    b, m = em.STFT_L2_distance(d["mixed_audio"], d["masks"], d["gt_bin_comps"].clone(), d["pred_mono"], d["gt_mono_comps"].clone())
    out["stft_l2_bin"], out["stft_l2_mono"] = b.numpy(), m.numpy()
    for ddppo in (False, True):
        torch.manual_seed(5 + int(ddppo))
        mod = mn.AcousticMem(use_ddppo=ddppo).eval()
        if not ddppo:       # non-trivial inference statistics
            g = torch.Generator().manual_seed(9)
            mod.cnn[1].weight.data = 0.5 + torch.rand(32, generator=g)
            mod.cnn[1].bias.data = 0.2 * torch.randn(32, generator=g)
            mod.cnn[1].running_mean.data = 0.2 * torch.randn(32, generator=g)
            mod.cnn[1].running_var.data = 0.5 + torch.rand(32, generator=g)
        with torch.no_grad():
            y = mod(d["pred_mono"], d["prev_mem"])
        tag = "ddppo" if ddppo else "bn"
        out[f"mem_{tag}_out"] = y.contiguous().numpy()
        for k, v in mod.state_dict().items():
            out[f"mem_{tag}_sd.{k}"] = v.numpy()
    w = waveform_case_inputs()
    metrics = ["si_sdr", "sd_sdr", "snr", "srr", "si_sdri", "sd_sdri", "snri", "si_sar"]
    q = em.compute_waveform_quality(w, metrics)
    out["wq_mono"] = np.array([q["mono"][k] for k in metrics], dtype=np.float64)
    out["wq_mem"] = np.array([q["monoFromMem"][k] for k in metrics], dtype=np.float64)
    out["wq_names"] = np.array(metrics)
    np.savez_compressed(os.path.join(GOLDEN, "postproc.npz"), **out)
    print("wrote postproc.npz", {k: v.shape for k, v in out.items()})


def main():
    if not os.path.isdir(REF_ROOT):
        sys.exit("needs /root/reference (build container only)")
    sim_train, sim_eval, ds_mod = load_reference_modules()
    os.makedirs(GOLDEN, exist_ok=True)
    gen_postproc()
    for name, K, lens, seed, norm in REF_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            monos, _ = write_case_files(tmp, K, lens, seed)
            first, sim = run_sim(sim_train.HabitatSimAudioEnabledTrain, tmp, K, norm)
            ev, _ = run_sim(sim_eval.HabitatSimAudioEnabledEval, tmp, K, norm, calls=1)
            dsout = run_dataset(ds_mod, tmp, K, norm, monos)
        out = dict(first[0])
        # second call (cache hits) and the eval twin must reproduce the first call exactly: stored as flags + asserted
        for key in first[0]:
            assert np.array_equal(first[0][key], first[1][key], equal_nan=True), ("cache-hit call differs", name, key)
            assert np.array_equal(first[0][key], ev[0][key], equal_nan=True), ("eval twin differs", name, key)
        for key in dsout[0]:
            assert np.array_equal(dsout[0][key], dsout[1][key]), ("dataset cache-hit call differs", name, key)
        assert sorted(sim._cached_mono_waveforms) == ["sound%d" % k for k in range(K)]
        out.update(dsout[0])
        out["lens"] = np.asarray(lens)
        out["seed_norm"] = np.asarray([seed, norm], dtype=np.float64)
        np.savez_compressed(os.path.join(GOLDEN, "ref_render_%s.npz" % name), **out)
        print("wrote ref_render_%s.npz" % name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
