"""CPU oracle: passive-separator U-Net (torch fp32, functional restatement).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Parity unpinned by the
reference's own tests; pinned against the unmodified reference module imported
from /root/reference (``oracle/gen_golden.py`` -> ``tests/golden/unet_*.npz``).

Follows ``audio_separation/rl/models/separator_cnn.py``:
``unet_conv :5-12``, ``unet_upconv :15-24``, ``PassiveSepEncCNN.forward :70-108``,
``PassiveSepDecCNN.forward :153-170``; and the loss of
``audio_separation/pretrain/passive/passive_trainer.py:269-286``.

The functions take plain ``state_dict``-style mappings whose keys are the
reference's own parameter names relative to the ``cnn`` Sequential
(``"cnn.{i}.0.weight"``, ``"cnn.{i}.1.{weight,bias,running_mean,running_var}"``,
decoder head ``"cnn.5.0.{weight,bias}"``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SLICE_FACTOR = 16      # separator_cnn.py:39, 122
BN_EPS = 1e-5          # nn.BatchNorm2d default
BN_MOMENTUM = 0.1
ENC_CH = [64, 128, 256, 512, 512]                 # separator_cnn.py:44-50
DEC_IN = [512, 1024, 512, 256, 128]               # separator_cnn.py:128-134


def enc_shapes(convert_bin2mono: bool = False):
    cin0 = 2 * SLICE_FACTOR + (0 if convert_bin2mono else 1)
    cins = [cin0] + ENC_CH[:-1]
    return [(co, ci, 4, 4) for ci, co in zip(cins, ENC_CH)]


def dec_shapes(convert_bin2mono: bool = False):
    n_out = SLICE_FACTOR * (1 if convert_bin2mono else 2)
    couts = [512, 256, 128, 64, n_out]
    return [(ci, co, 4, 4) for ci, co in zip(DEC_IN, couts)], (n_out, n_out, 1, 1)


def init_state(convert_bin2mono: bool = False, seed: int = 0, randomize_bn: bool = True):
    """Random encoder/decoder state dicts with the reference's shapes.

    Kaiming-normal conv weights as ``layer_init`` (separator_cnn.py:56-68,
    139-151); BatchNorm affine/running stats are randomised (instead of the
    1/0/0/1 defaults) when ``randomize_bn`` so that parity tests exercise them.
    """
    g = torch.Generator().manual_seed(seed)
    enc, dec = {}, {}
    for i, shp in enumerate(enc_shapes(convert_bin2mono)):
        fan_in = shp[1] * 16
        gain = torch.nn.init.calculate_gain("leaky_relu", 0.2)
        enc[f"cnn.{i}.0.weight"] = torch.randn(shp, generator=g) * (gain / fan_in ** 0.5)
        _bn(enc, f"cnn.{i}.1", shp[0], g, randomize_bn)
    shapes, head = dec_shapes(convert_bin2mono)
    for i, shp in enumerate(shapes):
        fan_in = shp[1] * 16                      # ConvTranspose2d: fan_in = weight.size(1) * k*k
        gain = torch.nn.init.calculate_gain("relu")
        dec[f"cnn.{i}.0.weight"] = torch.randn(shp, generator=g) * (gain / fan_in ** 0.5)
        _bn(dec, f"cnn.{i}.1", shp[1], g, randomize_bn)
    dec["cnn.5.0.weight"] = torch.randn(head, generator=g) * (2.0 / head[1]) ** 0.5
    dec["cnn.5.0.bias"] = (torch.randn(head[0], generator=g) * 0.1) if randomize_bn else torch.zeros(head[0])
    return enc, dec


def _bn(sd, prefix, c, g, randomize):
    if randomize:
        sd[prefix + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.2 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    else:
        sd[prefix + ".weight"] = torch.ones(c)
        sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _r(x, emulate_bf16):
    """Round to bf16 in the forward pass only (straight-through gradient).  With ``emulate_bf16`` the oracle rounds its GEMM
    operands and stored activations where the CUDA path does (fp32 accumulation kept), so that activation gates and
    BatchNorm statistics see the same values and a gradient comparison measures the backward arithmetic, not the
    ReLU-gate flips that any reduced-precision forward causes."""
    if not emulate_bf16:
        return x
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


def _batchnorm(x, sd, prefix, training):
    return F.batch_norm(x, None if training else sd[prefix + ".running_mean"],
                        None if training else sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"],
                        training=training, momentum=BN_MOMENTUM, eps=BN_EPS)


def encoder_input(mixed_bin_audio_mag, target_class=None, convert_bin2mono=False, mixed_audio=None):
    """separator_cnn.py:70-99: slice/reshape (+ class plane) -> ``[B, 32|33, 32, T]``."""
    if convert_bin2mono:
        x = mixed_bin_audio_mag * (torch.exp(mixed_audio) - 1)
        x = torch.log1p(torch.clamp(x, min=0))
    else:
        x = mixed_bin_audio_mag
    x = x.permute(0, 3, 1, 2)
    x = x.reshape(x.size(0), x.size(1), SLICE_FACTOR, -1, x.size(3))
    x = x.reshape(x.size(0), -1, x.size(3), x.size(4))
    if not convert_bin2mono:
        tc = target_class.reshape(x.size(0), 1, 1, 1).float() + 1
        x = torch.cat((x, tc.expand(-1, 1, x.size(2), x.size(3))), dim=1)
    return x


def encoder_forward(enc_sd, mixed_bin_audio_mag, target_class=None, convert_bin2mono=False,
                    mixed_audio=None, training=False, emulate_bf16=False):
    """-> ``(bottleneck [B,512], [skip 512@2x2, 256@4x4, 128@8x8, 64@16x16])``."""
    out = encoder_input(mixed_bin_audio_mag, target_class, convert_bin2mono, mixed_audio)
    if emulate_bf16:            # the class plane stays fp32 on the CUDA path (an epilogue term), the audio channels are bf16
        out = torch.cat((_r(out[:, :32], True), out[:, 32:]), dim=1)
    feats = []
    for i in range(5):
        w = enc_sd[f"cnn.{i}.0.weight"]
        if emulate_bf16 and i == 0 and w.size(1) == 33:
            w = torch.cat((_r(w[:, :32], True), w[:, 32:]), dim=1)
        else:
            w = _r(w, emulate_bf16)
        out = _r(F.conv2d(out, w, None, stride=2, padding=1), emulate_bf16)
        out = _batchnorm(out, enc_sd, f"cnn.{i}.1", training)
        out = _r(F.leaky_relu(out, 0.2), emulate_bf16)
        feats.append(out)
    return out.reshape(out.size(0), -1), feats[:-1][::-1]


def decoder_forward(dec_sd, bottleneck, skips, training=False, emulate_bf16=False):
    """-> mask ``[B, 512, T, C]`` (C = 2 binaural / 1 bin2mono); no sigmoid."""
    out = bottleneck.reshape(bottleneck.size(0), -1, 1, 1)
    for i in range(5):
        if i > 0:
            out = torch.cat((out, skips[i - 1]), dim=1)
        out = _r(F.conv_transpose2d(out, _r(dec_sd[f"cnn.{i}.0.weight"], emulate_bf16), None, stride=2, padding=1), emulate_bf16)
        out = _batchnorm(out, dec_sd, f"cnn.{i}.1", training)
        out = _r(F.relu(out), emulate_bf16)
    out = F.conv2d(out, _r(dec_sd["cnn.5.0.weight"], emulate_bf16), dec_sd["cnn.5.0.bias"])
    out = out.reshape(out.size(0), -1, SLICE_FACTOR, out.size(2), out.size(3))
    out = out.reshape(out.size(0), out.size(1), -1, out.size(4))
    return out.permute(0, 2, 3, 1)


def separator_forward(enc_sd, dec_sd, mixed_bin_audio_mag, target_class=None, convert_bin2mono=False,
                      mixed_audio=None, training=False, emulate_bf16=False):
    """``Policy.get_binSepMasks`` / ``convert_bin2mono`` (pretrain/passive/policy.py:61-71)."""
    b, s = encoder_forward(enc_sd, mixed_bin_audio_mag, target_class, convert_bin2mono, mixed_audio, training, emulate_bf16)
    return decoder_forward(dec_sd, b, s, training, emulate_bf16)


def supervised_losses(mixed_audio, pred_bin_masks, gt_bin_mag, pred_mono=None, gt_mono_mag=None):
    """passive_trainer.py:269-276: ``L1(mask * (exp(mix)-1), gt_bin)`` (+ mono L1)."""
    mix = torch.exp(mixed_audio) - 1
    bin_loss = F.l1_loss(pred_bin_masks * mix, gt_bin_mag)
    mono_loss = F.l1_loss(pred_mono, gt_mono_mag) if pred_mono is not None else None
    return bin_loss, mono_loss
