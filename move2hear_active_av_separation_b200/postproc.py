"""Consumers of the separator's masks on the rollout / evaluation step, drop-in by name:

* :func:`STFT_L2_distance`           - ``audio_separation/common/eval_metrics.py:306-366``
* :class:`AcousticMem`               - ``audio_separation/rl/models/memory_nets.py:5-69`` (same ctor, module tree, state_dict)
* :func:`compute_waveform_quality`   - ``audio_separation/common/eval_metrics.py:256-303`` (+ the nussl helpers :12-230)

All arithmetic runs in ``libm2h_b200.so`` (``csrc/m2h_post.cu``, ``csrc/render_stft.cuh::istft``); there is no CPU path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib

EPS = 1e-13      # eval_metrics.py:7
NAME_OF_ALL_QUALITY_METRICS = ['env', 'si_sdr', 'si_sir', 'si_sar', 'sd_sdr', 'snr', 'srr', 'si_sdri', 'sd_sdri', 'snri',
                               "si_siri", "si_sari", "sdr", "sir", "sar"]


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _cuda_f32(t, name, device=None):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.M2HError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if device is not None and t.device != device:
        raise _lib.M2HError(f"{name} is on {t.device}, expected {device}")
    return t if t.dtype == torch.float32 else t.float()


def _strides4(t):
    return (ctypes.c_int64 * 4)(*t.stride())


def STFT_L2_distance(mixed_audio, pred_binSepMasks, gt_bin_comps, pred_mono, gt_mono_comps, to_cpu: bool = True):
    """Same arguments and return value as the reference (two ``[B, 1]`` float32 tensors, on the CPU like the reference's
    ``.cpu()`` unless ``to_cpu=False``): STFT L2 distance of the predicted binaural / monaural spectrograms, carried by the
    ground-truth phases, from the ground truth."""
    lib = _lib.load()
    mixed = _cuda_f32(mixed_audio, "mixed_audio").contiguous()
    dev = mixed.device
    masks = _cuda_f32(pred_binSepMasks, "pred_binSepMasks", dev)
    gt_bin = _cuda_f32(gt_bin_comps, "gt_bin_comps", dev).contiguous()
    mono = _cuda_f32(pred_mono, "pred_mono", dev)
    gt_mono = _cuda_f32(gt_mono_comps, "gt_mono_comps", dev).contiguous()
    B, F, T = mixed.shape[:3]
    if F != 512 or mixed.shape[3] != 2 or tuple(masks.shape) != (B, 512, T, 2) or tuple(mono.shape[:3]) != (B, 512, T) \
            or tuple(gt_bin.shape[:3]) != (B, 512, T) or gt_bin.shape[3] < 4 or tuple(gt_mono.shape[:3]) != (B, 512, T) or gt_mono.shape[3] < 2:
        raise _lib.M2HError("STFT_L2_distance: expected mixed_audio / masks [B,512,T,2], gt_bin_comps [B,512,T,>=4], "
                            "pred_mono [B,512,T,1], gt_mono_comps [B,512,T,>=2]")
    out = torch.empty((2, B, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.m2h_stft_l2_distance(mixed.data_ptr(), masks.data_ptr(), _strides4(masks), gt_bin.data_ptr(), gt_bin.shape[3],
                                      mono.data_ptr(), _strides4(mono), gt_mono.data_ptr(), gt_mono.shape[3], B, T,
                                      out[0].data_ptr(), out[1].data_ptr(), _stream(dev))
    _lib.check(rc, "m2h_stft_l2_distance")
    if to_cpu:
        out = out.cpu()
    return out[0], out[1]


class AcousticMem(nn.Module):
    """Acoustic memory refiner (drop-in for memory_nets.py:5-69): the torch layers only hold the parameters, the forward is one
    fused CUDA kernel.  Inference statistics only (the rollout / evaluation call sites run under ``no_grad`` with the
    module in eval mode or, for ``use_ddppo=True``, without BatchNorm); a train-mode BatchNorm forward is refused."""

    def __init__(self, use_ddppo=False):
        super().__init__()
        self._slice_factor = 16
        n = self._slice_factor
        if use_ddppo:
            self.cnn = nn.Sequential(nn.Conv2d(n * 2, 32, kernel_size=3, padding=1, bias=False), nn.ReLU(inplace=True),
                                     nn.Conv2d(32, n, kernel_size=3, padding=1, bias=False))
        else:
            self.cnn = nn.Sequential(nn.Conv2d(n * 2, 32, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(32),
                                     nn.ReLU(inplace=True), nn.Conv2d(32, n, kernel_size=3, padding=1, bias=False))
        self.layer_init()

    def layer_init(self):
        for layer in self.cnn:
            if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)
            elif isinstance(layer, (nn.BatchNorm1d, nn.BatchNorm2d)):
                if layer.affine:
                    layer.weight.data.fill_(1)
                    layer.bias.data.zero_()

    def forward(self, pred_mono, prev_pred_monoFromMem_masked, out=None):
        lib = _lib.load()
        pm = _cuda_f32(pred_mono, "pred_mono")
        dev = pm.device
        pv = _cuda_f32(prev_pred_monoFromMem_masked, "prev_pred_monoFromMem_masked", dev)
        B = pm.size(0)
        if tuple(pm.shape) != (B, 512, 32, 1) or tuple(pv.shape) != (B, 512, 32, 1):
            raise _lib.M2HError(f"AcousticMem expects [B,512,32,1] spectrograms, got {tuple(pm.shape)} / {tuple(pv.shape)}")
        bn = self.cnn[1] if isinstance(self.cnn[1], nn.BatchNorm2d) else None
        if bn is not None and self.training:
            raise _lib.M2HError("AcousticMem on the CUDA path runs inference statistics only: call .eval() (the rollout and "
                                "evaluation call sites do)")
        w1, w2 = self.cnn[0].weight, self.cnn[-1].weight
        for name, t in (("cnn.0.weight", w1), ("cnn.-1.weight", w2)):
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.M2HError(f"AcousticMem parameter {name} must be a contiguous fp32 tensor on {dev}")
        if out is None:
            out = torch.empty((B, 512, 32, 1), dtype=torch.float32, device=dev)
        bnp = [None] * 4 if bn is None else [bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr()]
        with torch.cuda.device(dev):
            rc = lib.m2h_acoustic_mem_forward(pm.data_ptr(), _strides4(pm), pv.data_ptr(), _strides4(pv), w1.data_ptr(), w2.data_ptr(),
                                              *bnp, float(bn.eps) if bn is not None else 0.0, B, out.data_ptr(), _stream(dev))
        _lib.check(rc, "m2h_acoustic_mem_forward")
        return out


# ---------------------------------------------------------------------------------------------------------------------
# waveform-level quality (SI-SDR family)
# ---------------------------------------------------------------------------------------------------------------------

def _bss_from_moments(n, sr, se, srr, see, sre, compute_sir_sar=True):
    """scale_bss_eval_helper (eval_metrics.py:12-57) for ONE reference source, from the raw moments of the (uncentred)
    reference r and estimate e; preprocess() removes the means first (:180, :186)."""
    rr = srr - sr * sr / n          # sum (r - mean r)^2
    ee = see - se * se / n
    re = sre - sr * se / n
    alpha = re / rr
    snr = 10 * math.log10(rr / (ee - 2 * re + rr))
    signal = alpha * alpha * rr
    noise = ee - re * re / rr
    si_sdr = 10 * math.log10(signal / noise)
    srr_db = -10 * math.log10((1 - (1 / alpha)) ** 2)
    sd_sdr = snr + 10 * math.log10(alpha ** 2)
    si_sir = si_sar = float("nan")
    if compute_sir_sar:
        # one reference: the projection of the residual on the reference vanishes, b = EPS, e_interf = r * EPS
        b = EPS
        e_interf = b * b * rr
        e_artif = noise - 2 * b * 0.0 + e_interf          # |e_res - e_interf|^2 with <e_res, r> = 0 (the reference adds EPS element-wise: < 1e-9 relative)
        si_sir = 10 * math.log10(signal / e_interf)
        si_sar = 10 * math.log10(signal / e_artif)
    return si_sdr, si_sir, si_sar, sd_sdr, snr, srr_db


def waveform_metrics(reference: torch.Tensor, estimate: torch.Tensor, mixture: torch.Tensor, compute_sir_sar: bool = True):
    """``evaluate(true, estimated, mixed)`` (eval_metrics.py:199-230) for a batch: ``reference`` / ``estimate`` ``[B, L]``,
    ``mixture`` ``[B, C, L]`` CUDA fp32 -> list of B score dicts.  The O(L) work (five dot products per signal pair) runs
    in one kernel; the closed forms are evaluated in float64 on the 9 moments."""
    lib = _lib.load()
    ref = _cuda_f32(reference, "reference").contiguous()
    dev = ref.device
    est = _cuda_f32(estimate, "estimate", dev).contiguous()
    mix = _cuda_f32(mixture, "mixture", dev).contiguous()
    B, L = ref.shape
    if tuple(est.shape) != (B, L) or mix.dim() != 3 or mix.shape[0] != B or mix.shape[2] != L:
        raise _lib.M2HError("waveform_metrics: reference / estimate [B, L], mixture [B, C, L]")
    mom = torch.empty((B, 9), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.m2h_wave_moments(ref.data_ptr(), est.data_ptr(), mix.data_ptr(), mix.shape[1], B, L, mom.data_ptr(), _stream(dev))
    _lib.check(rc, "m2h_wave_moments")
    scores = []
    for n, sr, se, sm, srr, see, smm, sre, srm in mom.cpu().tolist():
        est_m = _bss_from_moments(n, sr, se, srr, see, sre, compute_sir_sar)
        mix_m = _bss_from_moments(n, sr, sm, srr, smm, srm, compute_sir_sar)
        si_sdr, si_sir, si_sar, sd_sdr, snr, srr_db = est_m
        scores.append({'si_sdr': si_sdr, 'si_sir': si_sir, 'si_sar': si_sar, 'sd_sdr': sd_sdr, 'snr': snr, 'srr': srr_db,
                       'si_sdri': si_sdr - mix_m[0], 'sd_sdri': sd_sdr - mix_m[3], 'snri': snr - mix_m[4],
                       'si_siri': si_sir - mix_m[1], 'si_sari': si_sar - mix_m[2]})
    return scores


def compute_waveform_quality(pred_n_gt_spects, eval_metrics_to_compute, rend=None):
    """Drop-in for ``eval_metrics.compute_waveform_quality`` (:256-303): inverse STFTs of the mixture, the ground truth and
    the two predictions (one batched ``m2h_istft`` launch for the three monaural signals, one for the binaural mixture),
    then the BSS metrics of 'mono' and 'monoFromMem'.  Inputs may be numpy arrays (as in the reference) or tensors."""
    from .render import get_renderer
    rend = rend or get_renderer()
    dev = rend.device

    def t(x):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x)
        return x.to(device=dev, dtype=torch.float32)

    mixed_mag, mixed_phase = t(pred_n_gt_spects["mixed_bin_audio_mag"])[:1], t(pred_n_gt_spects["mixed_bin_audio_phase"])[:1]
    gt_mag, gt_phase = t(pred_n_gt_spects["gt_mono_mag"])[:1, :, :, :1], t(pred_n_gt_spects["gt_mono_phase"])[:1, :, :, :1]
    pred_mono = t(pred_n_gt_spects["pred_mono"])[:1, :, :, :1]
    pred_mem = t(pred_n_gt_spects["pred_monoFromMem"])[:1, :, :, :1]
    mixed = rend.istft(mixed_mag.contiguous(), mixed_phase.contiguous(), length=16000)             # [1, 2, 16000]
    mags = torch.cat([gt_mag, pred_mono, pred_mem]).contiguous()                                  # the predictions ride the GT phase
    monos = rend.istft(mags, gt_phase.expand(3, -1, -1, -1).contiguous(), length=16000)[:, 0]      # [3, 16000]
    ref = monos[0:1].expand(2, -1).contiguous()
    scores = waveform_metrics(ref, monos[1:3].contiguous(), mixed.expand(2, -1, -1).contiguous())
    metrics = {"mono": {}, "monoFromMem": {}}
    for metric in eval_metrics_to_compute:
        assert metric in NAME_OF_ALL_QUALITY_METRICS, print("doesn't support computation of this metric")
        metrics["mono"][metric] = scores[0][metric]
        metrics["monoFromMem"][metric] = scores[1][metric]
    return metrics
