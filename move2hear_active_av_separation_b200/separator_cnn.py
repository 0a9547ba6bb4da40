"""Drop-in ``PassiveSepEncCNN`` / ``PassiveSepDecCNN`` backed by the sm_100a tcgen05 kernels.

Mirrors ``audio_separation/rl/models/separator_cnn.py`` (ctor :36-54 / :119-137, forward :70-108 / :153-170):
same constructor arguments, the same module tree (``cnn.{i}.0`` conv / deconv, ``cnn.{i}.1`` BatchNorm2d,
``cnn.5.0`` 1x1 head) so that ``state_dict`` keys, shapes, ``.parameters()``, ``.train()/.eval()``, ``.to()`` and
``requires_grad_`` behave exactly like the reference's modules and its checkpoints load unchanged; the same forward
signatures and return shapes.  The torch layers only *hold* the parameters - the arithmetic runs in
``libm2h_b200.so`` (``csrc/unet_gemm.cuh``) through the C ABI of ``include/m2h.h``.

Only the inference-statistics forward (``.eval()``: the RL rollout / update_sep / evaluation call sites,
ppo_trainer.py:300-307, 359-364, ppo.py:197-204) is implemented; calling a module in ``.train()`` mode raises - there
is no silent torch fallback.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib

SLICE_FACTOR = 16
_ENC_CH = (64, 128, 256, 512, 512)
_DEC_IO = ((512, 512), (1024, 256), (512, 128), (256, 64))


class _EncParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 5) for n in ("conv_w", "bn_w", "bn_b", "bn_mean", "bn_var")]


class _DecParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 5) for n in ("deconv_w", "bn_w", "bn_b", "bn_mean", "bn_var")] + \
               [("head_w", ctypes.c_void_p), ("head_b", ctypes.c_void_p)]


def _down_block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=False),
                         nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, True))


def _up_block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(True))


def _init_layers(seq: nn.Sequential, gain: float) -> None:
    for block in seq:
        for layer in block:
            if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(layer.weight, gain)
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)
            elif isinstance(layer, nn.BatchNorm2d):
                layer.weight.data.fill_(1)
                layer.bias.data.zero_()


def _f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _Workspace:
    """bf16 NHWC activations of one encoder/decoder pass (skip connections included)."""

    def __init__(self, B: int, device):
        lib = _lib.load()
        self.B = B
        self.nbytes = lib.m2h_unet_workspace_bytes(B)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
        offs = (ctypes.c_size_t * 7)()
        _lib.check(lib.m2h_unet_workspace_layout(B, offs), "m2h_unet_workspace_layout")
        self.offsets = list(offs)

    def view(self, idx: int, shape):
        """Logical (un-haloed) view of slot ``idx``: ``(B, H, W, C)`` feature map or the ``(B, 512)`` bottleneck."""
        n = 1
        full = tuple(shape) if len(shape) == 2 else (shape[0], shape[1] + 2, shape[2] + 2, shape[3])
        for s in full:
            n *= s
        t = self.buf[self.offsets[idx]: self.offsets[idx] + 2 * n].view(torch.bfloat16).view(full)
        return t if len(shape) == 2 else t[:, 1:-1, 1:-1, :]

    def clear_halos(self):
        lib = _lib.load()
        with torch.cuda.device(self.buf.device):
            rc = lib.m2h_unet_workspace_clear_halos(self.B, self.buf.data_ptr(), self.nbytes,
                                                    torch.cuda.current_stream(self.buf.device).cuda_stream)
        _lib.check(rc, "m2h_unet_workspace_clear_halos")

    def skips(self) -> List[torch.Tensor]:
        """NCHW-shaped views (bf16) of the encoder's skip features, deepest first like the reference."""
        B = self.B
        out = []
        for idx, (hw, c) in zip((4, 3, 2, 1), ((2, 512), (4, 256), (8, 128), (16, 64))):
            out.append(self.view(idx, (B, hw, hw, 2 * c))[..., c:].permute(0, 3, 1, 2))
        return out


class _PackedModule(nn.Module):
    """Common parameter-packing cache: re-pack when any parameter/buffer changed (load_state_dict, optimiser step)."""

    def _signature(self):
        sig = []
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append((t.data_ptr(), t._version, t.device.index, t.dtype))
        return tuple(sig)

    def _require_cuda_eval(self, device):
        if device.type != "cuda":
            raise _lib.M2HError("the B200 U-Net has no CPU path: move the module and its inputs to a CUDA device")
        if self.training:
            raise _lib.M2HError("train-mode (batch-statistics) forward/backward is not implemented by the CUDA path; "
                                "call .eval() (RL rollout / update_sep / evaluation call sites)")


class PassiveSepEncCNN(_PackedModule):
    r"""U-Net encoder for passive separation (drop-in for separator_cnn.py:27-108).

    Args:
        convert_bin2mono: creates the encoder of the binaural-to-monaural converter if set to True
    """

    def __init__(self, convert_bin2mono: bool = False):
        super().__init__()
        self._convert_bin2mono = convert_bin2mono
        self._slice_factor = SLICE_FACTOR
        self._n_input_audio = 2 * SLICE_FACTOR + (0 if convert_bin2mono else 1)
        chans = (self._n_input_audio,) + _ENC_CH
        self.cnn = nn.Sequential(*[_down_block(chans[i], chans[i + 1]) for i in range(5)])
        self.layer_init()
        self._packed = None
        self._packed_sig = None

    def layer_init(self):
        _init_layers(self.cnn, nn.init.calculate_gain("leaky_relu", 0.2))

    def _pack(self, device):
        sig = self._signature()
        if self._packed is not None and self._packed_sig == sig:
            return self._packed
        lib = _lib.load()
        keep, p = [], _EncParams()
        for i in range(5):
            conv, bn = self.cnn[i][0], self.cnn[i][1]
            for name, t in (("conv_w", conv.weight), ("bn_w", bn.weight), ("bn_b", bn.bias),
                            ("bn_mean", bn.running_mean), ("bn_var", bn.running_var)):
                t = _f32(t)
                keep.append(t)
                getattr(p, name)[i] = t.data_ptr()
        packed = torch.empty(lib.m2h_unet_enc_packed_bytes(int(self._convert_bin2mono)), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            rc = lib.m2h_unet_pack_encoder(ctypes.byref(p), int(self._convert_bin2mono), float(self.cnn[0][1].eps),
                                           packed.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_pack_encoder")
        self._packed, self._packed_sig = packed, sig
        return packed

    def forward(self, observations, mixed_audio=None):
        lib = _lib.load()
        if self._convert_bin2mono:
            assert mixed_audio is not None
            masks = observations                       # predicted binaural ratio masks [B,512,32,2] (any strides)
            obs = mixed_audio
        else:
            obs = observations["mixed_bin_audio_mag"]
            masks = None
        device = obs.device
        self._require_cuda_eval(device)
        obs = _f32(obs)
        B = obs.size(0)
        if tuple(obs.shape[1:]) != (512, 32, 2):
            raise _lib.M2HError(f"expected a [B,512,32,2] spectrogram, got {tuple(obs.shape)}")
        packed = self._pack(device)
        ws = _Workspace(B, device)
        bott = torch.empty((B, 512), dtype=torch.float32, device=device)
        cls_ptr = mask_ptr = strides = None
        keep = []
        if self._convert_bin2mono:
            masks = masks.detach()
            if masks.dtype != torch.float32:
                masks = masks.float()
            keep.append(masks)
            mask_ptr = masks.data_ptr()
            strides = (ctypes.c_int64 * 4)(*masks.stride())
        else:
            cls = (observations["target_class"].detach().reshape(B).float() + 1).contiguous()
            keep.append(cls)
            cls_ptr = cls.data_ptr()
        with torch.cuda.device(device):
            rc = lib.m2h_unet_encoder_forward(packed.data_ptr(), int(self._convert_bin2mono), obs.data_ptr(), cls_ptr,
                                              mask_ptr, strides, B, ws.buf.data_ptr(), ws.nbytes, bott.data_ptr(),
                                              torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_encoder_forward")
        skips = ws.skips()
        bott._m2h_ws = ws                              # lets the partner decoder reuse the activations in place
        return bott, skips


class PassiveSepDecCNN(_PackedModule):
    r"""U-Net decoder for passive separation (drop-in for separator_cnn.py:111-170).  No sigmoid.

    Args:
        convert_bin2mono: creates the decoder of the binaural-to-monaural converter if set to True
    """

    def __init__(self, convert_bin2mono: bool = False):
        super().__init__()
        self._convert_bin2mono = convert_bin2mono
        self._slice_factor = SLICE_FACTOR
        self._n_out_audio = SLICE_FACTOR * (1 if convert_bin2mono else 2)
        blocks = [_up_block(ci, co) for ci, co in _DEC_IO] + [_up_block(128, self._n_out_audio)]
        blocks.append(nn.Sequential(nn.Conv2d(self._n_out_audio, self._n_out_audio, kernel_size=1)))
        self.cnn = nn.Sequential(*blocks)
        self.layer_init()
        self._packed = None
        self._packed_sig = None

    def layer_init(self):
        _init_layers(self.cnn, nn.init.calculate_gain("relu"))

    def _pack(self, device):
        sig = self._signature()
        if self._packed is not None and self._packed_sig == sig:
            return self._packed
        lib = _lib.load()
        keep, p = [], _DecParams()
        for i in range(5):
            conv, bn = self.cnn[i][0], self.cnn[i][1]
            for name, t in (("deconv_w", conv.weight), ("bn_w", bn.weight), ("bn_b", bn.bias),
                            ("bn_mean", bn.running_mean), ("bn_var", bn.running_var)):
                t = _f32(t)
                keep.append(t)
                getattr(p, name)[i] = t.data_ptr()
        hw, hb = _f32(self.cnn[5][0].weight), _f32(self.cnn[5][0].bias)
        p.head_w, p.head_b = hw.data_ptr(), hb.data_ptr()
        packed = torch.empty(lib.m2h_unet_dec_packed_bytes(int(self._convert_bin2mono)), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            rc = lib.m2h_unet_pack_decoder(ctypes.byref(p), int(self._convert_bin2mono), float(self.cnn[0][1].eps),
                                           packed.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_pack_decoder")
        self._packed, self._packed_sig = packed, sig
        return packed

    def forward(self, bottleneck_feats, lst_skip_feats):
        lib = _lib.load()
        device = bottleneck_feats.device
        self._require_cuda_eval(device)
        B = bottleneck_feats.size(0)
        ws: Optional[_Workspace] = getattr(bottleneck_feats, "_m2h_ws", None)
        if ws is None or ws.B != B or ws.buf.device != device:
            # features did not come straight from the partner encoder: stage them (layout/dtype copies only)
            ws = _Workspace(B, device)
            ws.clear_halos()
            ws.view(5, (B, 512)).copy_(bottleneck_feats.reshape(B, 512))
            for idx, feat, (hw, c) in zip((4, 3, 2, 1), lst_skip_feats, ((2, 512), (4, 256), (8, 128), (16, 64))):
                ws.view(idx, (B, hw, hw, 2 * c))[..., c:].copy_(feat.permute(0, 2, 3, 1))
        packed = self._pack(device)
        C = self._n_out_audio // SLICE_FACTOR
        out = torch.empty((B, C, 512, 32), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            rc = lib.m2h_unet_decoder_forward(packed.data_ptr(), int(self._convert_bin2mono), B, ws.buf.data_ptr(),
                                              ws.nbytes, out.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_decoder_forward")
        # [BATCH x HEIGHT x WIDTH x CHANNEL] view with the reference's strides (C*16384, 32, 1, 16384)
        return out.permute(0, 2, 3, 1)


def smoke_check(obs: torch.Tensor) -> None:
    """Tiny U-Net forward on ``obs`` ([B,512,32,2] log1p magnitudes) checked against the oracle (used by smoke())."""
    from oracle import unet as U
    enc_sd, dec_sd = U.init_state(False, seed=3)
    enc, dec = PassiveSepEncCNN().cuda().eval(), PassiveSepDecCNN().cuda().eval()
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)
    tc = torch.arange(obs.size(0), device=obs.device).reshape(-1, 1) % 11
    with torch.no_grad():
        mask = dec(*enc({"mixed_bin_audio_mag": obs, "target_class": tc}))
        want = U.separator_forward(enc_sd, dec_sd, obs.cpu(), tc.cpu())
    err = (mask.cpu() - want).norm() / want.norm()
    assert float(err) < 2e-2, float(err)
