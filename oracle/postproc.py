"""CPU oracle for the consumers of the masks.  TEST INFRASTRUCTURE ONLY.

Restates, term by term:

* ``STFT_L2_distance`` - audio_separation/common/eval_metrics.py:306-366 (torch, fp32);
* ``AcousticMem.forward`` - audio_separation/rl/models/memory_nets.py:40-69 (functional form on a state_dict);
* ``scale_bss_eval_helper`` / ``scale_bss_eval`` / ``preprocess`` / ``evaluate`` - eval_metrics.py:12-230 (numpy), and
  ``compute_waveform_quality`` (:256-303) on top of :func:`oracle.render.istft_eval`.

Pinned against the unmodified reference functions by ``oracle/gen_golden.py::gen_postproc`` ->
``tests/golden/postproc.npz`` (``eval_metrics.py`` imports under a ``librosa`` stub; ``memory_nets.py`` is torch-only).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import render as R

EPS = 1e-13


def _real_imag(mag, phase):
    x = torch.cat([(mag * torch.cos(phase)).unsqueeze(0), (mag * torch.sin(phase)).unsqueeze(0)], dim=0)
    return x.permute(1, 0, 2, 3).unsqueeze(1).contiguous().view(x.size(1), 1, -1)


def stft_l2_distance(mixed_audio, pred_masks, gt_bin_comps, pred_mono, gt_mono_comps):
    """eval_metrics.py:306-366 -> (bin [B,1], mono [B,1])."""
    gt_l, ph_l = gt_bin_comps[..., 0], gt_bin_comps[..., 1]
    gt_r, ph_r = gt_bin_comps[..., 2], gt_bin_comps[..., 3]
    pred_bin = (torch.exp(mixed_audio) - 1) * pred_masks
    d_l = torch.mean(torch.pow(_real_imag(gt_l, ph_l) - _real_imag(pred_bin[..., 0], ph_l), 2), dim=2)
    d_r = torch.mean(torch.pow(_real_imag(gt_r, ph_r) - _real_imag(pred_bin[..., 1], ph_r), 2), dim=2)
    gm, pm = gt_mono_comps[..., 0], gt_mono_comps[..., 1]
    d_m = torch.mean(torch.pow(_real_imag(gm, pm) - _real_imag(pred_mono[..., 0], pm), 2), dim=2)
    return d_l + d_r, d_m


def acoustic_mem_forward(sd, pred_mono, prev, use_ddppo=False, slice_factor=16):
    """memory_nets.py:40-69 with the module in eval mode; ``sd`` holds ``cnn.*`` keys of the reference module."""
    def slice_(x):
        x = x.permute(0, 3, 1, 2)
        x = x.view(x.size(0), x.size(1), slice_factor, -1, x.size(3))
        return x.reshape(x.size(0), -1, x.size(3), x.size(4))
    out = torch.cat((slice_(pred_mono), slice_(prev)), dim=1)
    out = F.conv2d(out, sd["cnn.0.weight"], None, padding=1)
    if use_ddppo:
        last = "cnn.2.weight"
    else:
        out = F.batch_norm(out, sd["cnn.1.running_mean"], sd["cnn.1.running_var"], sd["cnn.1.weight"], sd["cnn.1.bias"], training=False, eps=1e-5)
        last = "cnn.3.weight"
    out = F.conv2d(F.relu(out), sd[last], None, padding=1)
    out = out.view(out.size(0), -1, slice_factor, out.size(2), out.size(3))
    out = out.reshape(out.size(0), out.size(1), -1, out.size(4))
    return out.permute(0, 2, 3, 1)


def scale_bss_eval_helper(references, estimate, idx, compute_sir_sar=True):
    """eval_metrics.py:12-57."""
    source = references[..., idx]
    source_energy = (source ** 2).sum()
    alpha = source @ estimate / source_energy
    e_true = source
    e_res = estimate - e_true
    snr = 10 * np.log10((e_true ** 2).sum() / (e_res ** 2).sum())
    e_true = source * alpha
    e_res = estimate - e_true
    signal = (e_true ** 2).sum()
    noise = (e_res ** 2).sum()
    si_sdr = 10 * np.log10(signal / noise)
    srr = -10 * np.log10((1 - (1 / alpha)) ** 2)
    sd_sdr = snr + 10 * np.log10(alpha ** 2)
    si_sir = si_sar = np.nan
    if compute_sir_sar:
        references_projection = references.T @ references
        references_onto_residual = np.dot(references.transpose(), e_res)
        b = np.linalg.solve(references_projection, references_onto_residual) + EPS
        e_interf = np.dot(references, b)
        e_artif = e_res - e_interf + EPS
        si_sir = 10 * np.log10(signal / (e_interf ** 2).sum())
        si_sar = 10 * np.log10(signal / (e_artif ** 2).sum())
    return si_sdr, si_sir, si_sar, sd_sdr, snr, srr


def evaluate(true_signal, estimated_signal, mixed_signal, compute_sir_sar=True):
    """eval_metrics.py:167-230 (preprocess + evaluate_helper + scale_bss_eval)."""
    references = np.stack([x for x in true_signal], axis=-1).transpose(1, 0, 2)
    references = references - references.mean(axis=0)
    estimates = np.stack([x for x in estimated_signal], axis=-1).transpose(1, 0, 2)
    estimates = estimates - estimates.mean(axis=0)
    assert len(mixed_signal) == 1
    mixture = mixed_signal[0].transpose(1, 0) - mixed_signal[0].transpose(1, 0).mean(axis=0)
    mixture = np.mean(mixture, axis=1, keepdims=True)
    refs, est, mix = references[..., 0, :], estimates[..., 0, 0], mixture[..., 0]
    m = scale_bss_eval_helper(refs, est, 0, compute_sir_sar)
    mm = scale_bss_eval_helper(refs, mix, 0, compute_sir_sar)
    return {'si_sdr': m[0], 'si_sir': m[1], 'si_sar': m[2], 'sd_sdr': m[3], 'snr': m[4], 'srr': m[5],
            'si_sdri': m[0] - mm[0], 'sd_sdri': m[3] - mm[3], 'snri': m[4] - mm[4], 'si_siri': m[1] - mm[1], 'si_sari': m[2] - mm[2]}


def compute_waveform_quality(spects, metrics_to_compute):
    """eval_metrics.py:256-303 on the oracle's librosa-0.8.0 istft restatement."""
    mixed = [np.array(R.istft_eval(spects["mixed_bin_audio_mag"][0, :, :, 0], spects["mixed_bin_audio_phase"][0, :, :, 0],
                                   spects["mixed_bin_audio_mag"][0, :, :, 1], spects["mixed_bin_audio_phase"][0, :, :, 1]))]
    gt = [np.array(R.istft_eval(spects["gt_mono_mag"][0, :, :, 0], spects["gt_mono_phase"][0, :, :, 0]))]
    pm = [np.array(R.istft_eval(spects["pred_mono"][0, :, :, 0], spects["gt_mono_phase"][0, :, :, 0]))]
    pmm = [np.array(R.istft_eval(spects["pred_monoFromMem"][0, :, :, 0], spects["gt_mono_phase"][0, :, :, 0]))]
    a, b = evaluate(gt, pm, mixed), evaluate(gt, pmm, mixed)
    return {"mono": {k: a[k] for k in metrics_to_compute}, "monoFromMem": {k: b[k] for k in metrics_to_compute}}
