"""Generate the committed golden vectors under ``tests/golden/``.

Run in the BUILD container only (needs /root/reference):

    python -m oracle.gen_golden

* ``unet_*.npz``   - outputs of the UNMODIFIED reference module
  ``/root/reference/audio_separation/rl/models/separator_cnn.py`` (imported by
  path; torch-only) on seeded inputs with the seeded weights of
  ``oracle.unet.init_state``.  They pin ``oracle/unet.py`` and, through it, the
  CUDA U-Net.
* ``render_*.npz`` - outputs of the reference's own dependency call
  ``scipy.signal.fftconvolve`` composed exactly as
  ``simulator_train.py:386-486`` / ``dataset.py:162-228`` do (through
  ``oracle/render.py``), plus two independent cross-routes (float64
  ``np.convolve``; ``torch.stft`` float64) so that the restated librosa STFT is
  itself pinned.

Inputs are regenerated from seeds at test time (numpy ``default_rng`` and torch
CPU generators are stable), so only outputs are stored.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

from . import render as R
from . import unet as U

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = "/root/reference/audio_separation/rl/models/separator_cnn.py"

# (name, K, L list, seed) - small enough to commit, wide enough to cover the edge cases
RENDER_CASES = [
    ("k2_L16000", 2, [16000, 16000], 11),
    ("k2_ragged", 2, [7999, 4000], 12),
    ("k2_short", 2, [1, 2], 13),
    ("k2_zero_rir", 2, [0, 5000], 14),
    ("k3_long", 3, [20000, 32000, 16001], 15),
    ("k4_48000", 4, [48000, 40000, 8000, 16000], 16),
]


def render_case_inputs(K, lens, seed):
    monos = R.synth_monos(1, K, seed=seed)[0]
    Lmax = max(max(lens), 1)
    rirs_planar, _ = R.synth_rirs(1, K, L=Lmax, seed=seed + 1000)
    rirs_planar = rirs_planar[0]
    rng = np.random.default_rng(seed + 2000)
    rirs = []
    for k, lk in enumerate(lens):
        # re-draw every RIR at its own length so each has its own decay envelope
        t = np.arange(max(lk, 1), dtype=np.float32)
        h = rng.standard_normal((2, max(lk, 1))).astype(np.float32) * np.exp(-t / (max(lk, 1) / 6.0)).astype(np.float32) * np.float32(0.05)
        rirs_planar[k] = 0
        rirs_planar[k, :, :lk] = h[:, :lk]
        rirs.append(R.rir_wav_layout(rirs_planar[k], lk))
    return monos, rirs, rirs_planar, np.asarray(lens, dtype=np.int32)


def unet_case_inputs(B, seed, T=32):
    g = torch.Generator().manual_seed(seed)
    # log1p-magnitude-like, non-negative, with a realistic range
    obs = torch.log1p(torch.rand(B, 512, T, 2, generator=g) * 3.0)
    tc = torch.randint(0, 11, (B, 1), generator=g)
    return obs, tc


def _load_reference():
    spec = importlib.util.spec_from_file_location("ref_separator_cnn", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_unet():
    ref = _load_reference()
    for bin2mono in (False, True):
        enc_sd, dec_sd = U.init_state(bin2mono, seed=3 + int(bin2mono))
        enc = ref.PassiveSepEncCNN(convert_bin2mono=bin2mono)
        dec = ref.PassiveSepDecCNN(convert_bin2mono=bin2mono)
        enc.load_state_dict(enc_sd)
        dec.load_state_dict(dec_sd)
        obs, tc = unet_case_inputs(3, seed=5)
        out = {}
        for mode in ("eval", "train"):
            enc.train(mode == "train")
            dec.train(mode == "train")
            enc.load_state_dict(enc_sd)
            dec.load_state_dict(dec_sd)     # reset running stats
            with torch.no_grad():
                if bin2mono:
                    g = torch.Generator().manual_seed(9)
                    masks = torch.randn(3, 512, 32, 2, generator=g) * 0.5 + 0.3
                    bott, skips = enc(masks, mixed_audio=obs)
                    out["in_masks"] = masks.numpy()
                else:
                    bott, skips = enc({"mixed_bin_audio_mag": obs, "target_class": tc})
                mask = dec(bott, skips)
            out[f"{mode}_bottleneck"] = bott.numpy()
            out[f"{mode}_skip3"] = skips[3].numpy().astype(np.float16)
            out[f"{mode}_mask"] = mask.contiguous().numpy()
            out[f"{mode}_mask_strides"] = np.array(mask.stride())
            if mode == "train":
                out["train_enc0_running_mean"] = enc.state_dict()["cnn.0.1.running_mean"].numpy()
                out["train_dec4_running_var"] = dec.state_dict()["cnn.4.1.running_var"].numpy()
        out["weight_checksum"] = np.array([float(sum(v.double().sum() for v in enc_sd.values())),
                                           float(sum(v.double().sum() for v in dec_sd.values()))])
        name = "unet_bin2mono.npz" if bin2mono else "unet_binsep.npz"
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print("wrote", name, {k: v.shape for k, v in out.items()})


def gen_render():
    for name, K, lens, seed in RENDER_CASES:
        monos, rirs, _, _ = render_case_inputs(K, lens, seed)
        sim = R.render_sim(list(monos), rirs)
        ds = R.render_dataset(list(monos), rirs, target_class=3)
        # independent route for the convolution: exact float64 direct convolution
        k0 = 0 if lens[0] > 0 else 1
        full = np.convolve(monos[k0].astype(np.float64), rirs[k0][:, 0].astype(np.float64))
        c = (lens[k0] - 1) // 2
        direct = full[c:c + monos.shape[1]]
        out = {
            "sim_mixed_mag_f16": sim["mixed_bin_audio_mag"],
            "sim_mixed_phase_f16": sim["mixed_bin_audio_phase"],
            "sim_gt_bin_f16": np.stack(sim["gt_bin_audio_components"]),
            "sim_gt_mono_f16": np.stack(sim["gt_mono_audio_components"]),
            "mixed_wave_i32": np.round(sim["_mixed_wave"].astype(np.float64) * 32768 * K).astype(np.int32),
            "source_waves_i16": np.round(sim["_source_waves"] * 32768).astype(np.int16),
            "ds_mixed_logmag_f32": ds[0],
            "ds_gt_bin_mag_f32": ds[1][0],
            "ds_gt_mono_mag_f32": ds[2][0],
            "direct_conv_src_ear0_f64": direct,
            "direct_conv_src": np.array([k0]),
        }
        np.savez_compressed(os.path.join(GOLDEN, f"render_{name}.npz"), **out)
        print("wrote render_%s.npz" % name)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if not os.path.exists(REF):
        sys.exit("needs /root/reference (build container only)")
    gen_unet()
    gen_render()


if __name__ == "__main__":
    main()
