"""``torch.library`` registration of the hot path: the ``m2h::`` operator namespace.

``BASELINE.json::north_star`` asks for a thin C-ABI extension *registered with PyTorch*.  The kernels live behind the C ABI
of ``include/m2h.h`` (``libm2h_b200.so``, loaded with ctypes); this module declares one PyTorch operator per entry-point
family and registers its implementation for the **CUDA dispatch key only** - calling an ``m2h::`` operator on CPU tensors
fails in the dispatcher (``NotImplementedError: ... 'CPU' backend``): there is no CPU fallback to fall into.  Fake
(meta) implementations give the operators shapes / dtypes for tracing tools; the operators themselves are opaque launches
on the current CUDA stream, so they compose with CUDA graphs (``graphs.GraphedSeparation``).

    torch.ops.m2h.rir_convolve(mono, rir, rir_len)                      -> waves int16 [B,K,2,N]
    torch.ops.m2h.stft_waves(waves, n_gt, sim_dtypes, want_phase, raw)  -> (mix_mag, mix_phase, gt)     (empty = not requested)
    torch.ops.m2h.render_mixture(mono, rir, rir_len, n_gt, sim_dtypes, want_phase) -> (mix_mag, mix_phase, gt, waves)
    torch.ops.m2h.istft(mag, phase, length)                             -> waveforms [B,C,length]
    torch.ops.m2h.unet_forward(obs, target_class, enc_packed, dec_packed)            -> mask [B,2,512,32]  (view it .permute(0,2,3,1))
    torch.ops.m2h.bin2mono_forward(masks, obs, enc_packed, dec_packed)               -> mono [B,1,512,32]
    torch.ops.m2h.stft_l2_distance(mixed_audio, masks, gt_bin, pred_mono, gt_mono)   -> (bin [B,1], mono [B,1])
    torch.ops.m2h.acoustic_mem(pred_mono, prev, conv1_w, conv2_w, bn_w, bn_b, bn_mean, bn_var, bn_eps) -> [B,512,32,1]

Reference call sites: simulator_train.py:386-486, dataset.py:162-228 (render); separator_cnn.py:70-108, 153-170 (U-Net);
eval_metrics.py:232-252, 306-366; memory_nets.py:40-69.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_LIB = torch.library.Library("m2h", "DEF")
_LIB.define("rir_convolve(Tensor mono, Tensor rir, Tensor rir_len) -> Tensor")
_LIB.define("stft_waves(Tensor waves, int n_gt, bool sim_dtypes, bool want_phase, bool raw_mag) -> (Tensor, Tensor, Tensor)")
_LIB.define("render_mixture(Tensor mono, Tensor rir, Tensor rir_len, int n_gt, bool sim_dtypes, bool want_phase) -> (Tensor, Tensor, Tensor, Tensor)")
_LIB.define("stft_waves_x0(Tensor waves) -> Tensor")
_LIB.define("unet_forward_x0(Tensor x0, Tensor target_class, Tensor enc_packed, Tensor dec_packed) -> Tensor")
_LIB.define("istft(Tensor mag, Tensor phase, int length) -> Tensor")
_LIB.define("unet_forward(Tensor obs, Tensor target_class, Tensor enc_packed, Tensor dec_packed) -> Tensor")
_LIB.define("bin2mono_forward(Tensor masks, Tensor obs, Tensor enc_packed, Tensor dec_packed) -> Tensor")
_LIB.define("stft_l2_distance(Tensor mixed_audio, Tensor masks, Tensor gt_bin_comps, Tensor pred_mono, Tensor gt_mono_comps) -> (Tensor, Tensor)")
_LIB.define("acoustic_mem(Tensor pred_mono, Tensor prev, Tensor conv1_w, Tensor conv2_w, Tensor? bn_w, Tensor? bn_b, Tensor? bn_mean, Tensor? bn_var, float bn_eps) -> Tensor")


def _renderer(t):
    from .render import get_renderer
    return get_renderer(t.device)


def _empty(t):
    return torch.empty(0, device=t.device)


def _rir_convolve(mono, rir, rir_len):
    return _renderer(mono).rir_convolve(mono, rir, rir_len)


def _stft_waves(waves, n_gt, sim_dtypes, want_phase, raw_mag):
    mix, phase, gt = _renderer(waves).stft_waves(waves, n_gt, sim_dtypes, want_phase, raw_mag)
    return mix, phase if phase is not None else _empty(waves), gt if gt is not None else _empty(waves)


def _render_mixture(mono, rir, rir_len, n_gt, sim_dtypes, want_phase):
    waves = _rir_convolve(mono, rir, rir_len)
    mix, phase, gt = _stft_waves(waves, n_gt, sim_dtypes, want_phase, False)
    return mix, phase, gt, waves


def _stft_waves_x0(waves):
    """int16 ``[B,K,2,N]`` (32 frames) -> bf16 ``[B,18,18,128]``: log1p |STFT| of the mixture in the separator's first-layer layout."""
    return _renderer(waves).stft_waves_x0(waves).x0


def _istft(mag, phase, length):
    return _renderer(mag).istft(mag, phase, length)


_ws_cache = {}


def _workspace(B, device):
    from .separator_cnn import _Workspace
    key = (B, device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        if len(_ws_cache) > 8:
            _ws_cache.clear()
        ws = _ws_cache[key] = _Workspace(B, device)
    return ws


def _separator(obs, enc_packed, dec_packed, bin2mono, cls=None, masks=None, x0=None):
    lib = _lib.load()
    dev = (obs if x0 is None else x0).device
    if x0 is None:
        obs = obs.contiguous().float()
        B = obs.size(0)
        if tuple(obs.shape[1:]) != (512, 32, 2):
            raise _lib.M2HError(f"expected a [B,512,32,2] spectrogram, got {tuple(obs.shape)}")
    else:
        B = x0.size(0)
        if x0.dtype != torch.bfloat16 or tuple(x0.shape[1:]) != (18, 18, 128) or not x0.is_contiguous():
            raise _lib.M2HError(f"expected a contiguous bf16 [B,18,18,128] first-layer operand, got {x0.dtype} {tuple(x0.shape)}")
    for name, t in (("enc_packed", enc_packed), ("dec_packed", dec_packed)):
        if t.device != dev or t.dtype != torch.uint8:
            raise _lib.M2HError(f"{name} must be the uint8 buffer of m2h_unet_pack_* on {dev}")
    ws = _workspace(B, dev)
    bott = torch.empty((B, 512), dtype=torch.float32, device=dev)
    out = torch.empty((B, 1 if bin2mono else 2, 512, 32), dtype=torch.float32, device=dev)
    cls_ptr = mask_ptr = strides = None
    if bin2mono:
        masks = masks.float()
        mask_ptr, strides = masks.data_ptr(), (ctypes.c_int64 * 4)(*masks.stride())
    elif x0 is not None:
        cls = cls.reshape(B).to(torch.int64).contiguous()
        cls_ptr = cls.data_ptr()
    else:
        cls = (cls.reshape(B).float() + 1).contiguous()
        cls_ptr = cls.data_ptr()
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        if x0 is not None:
            _lib.check(lib.m2h_unet_encoder_forward_x0(enc_packed.data_ptr(), x0.data_ptr(), cls_ptr, B, ws.buf.data_ptr(), ws.nbytes,
                                                       bott.data_ptr(), st), "m2h_unet_encoder_forward_x0")
        else:
            _lib.check(lib.m2h_unet_encoder_forward(enc_packed.data_ptr(), int(bin2mono), obs.data_ptr(), cls_ptr, mask_ptr, strides, B,
                                                    ws.buf.data_ptr(), ws.nbytes, bott.data_ptr(), st), "m2h_unet_encoder_forward")
        _lib.check(lib.m2h_unet_decoder_forward(dec_packed.data_ptr(), int(bin2mono), B, ws.buf.data_ptr(), ws.nbytes, out.data_ptr(), st),
                   "m2h_unet_decoder_forward")
    return out


def _unet_forward(obs, target_class, enc_packed, dec_packed):
    return _separator(obs, enc_packed, dec_packed, False, cls=target_class)


def _unet_forward_x0(x0, target_class, enc_packed, dec_packed):
    return _separator(None, enc_packed, dec_packed, False, cls=target_class, x0=x0)


def _bin2mono_forward(masks, obs, enc_packed, dec_packed):
    return _separator(obs, enc_packed, dec_packed, True, masks=masks)


def _stft_l2(mixed_audio, masks, gt_bin_comps, pred_mono, gt_mono_comps):
    from .postproc import STFT_L2_distance
    return STFT_L2_distance(mixed_audio, masks, gt_bin_comps, pred_mono, gt_mono_comps, to_cpu=False)


def _acoustic_mem(pred_mono, prev, conv1_w, conv2_w, bn_w, bn_b, bn_mean, bn_var, bn_eps):
    lib = _lib.load()
    dev = pred_mono.device
    pm, pv = pred_mono.float(), prev.float()
    B = pm.size(0)
    out = torch.empty((B, 512, 32, 1), dtype=torch.float32, device=dev)
    bnp = [t.data_ptr() if t is not None else None for t in (bn_w, bn_b, bn_mean, bn_var)]
    with torch.cuda.device(dev):
        rc = lib.m2h_acoustic_mem_forward(pm.data_ptr(), (ctypes.c_int64 * 4)(*pm.stride()), pv.data_ptr(), (ctypes.c_int64 * 4)(*pv.stride()),
                                          conv1_w.data_ptr(), conv2_w.data_ptr(), *bnp, float(bn_eps), B, out.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "m2h_acoustic_mem_forward")
    return out


for _name, _fn in (("rir_convolve", _rir_convolve), ("stft_waves", _stft_waves), ("render_mixture", _render_mixture), ("istft", _istft),
                   ("unet_forward", _unet_forward), ("bin2mono_forward", _bin2mono_forward), ("stft_l2_distance", _stft_l2),
                   ("acoustic_mem", _acoustic_mem), ("stft_waves_x0", _stft_waves_x0), ("unet_forward_x0", _unet_forward_x0)):
    _LIB.impl(_name, _fn, "CUDA")


# ---- fake (meta) implementations: shapes and dtypes only ---------------------------------------------------------------
def _T(n):
    return 1 + (n - 1) // 512


@torch.library.register_fake("m2h::stft_waves_x0")
def _(waves):
    return waves.new_empty((waves.shape[0], 18, 18, 128), dtype=torch.bfloat16)


@torch.library.register_fake("m2h::unet_forward_x0")
def _(x0, target_class, enc_packed, dec_packed):
    return x0.new_empty((x0.shape[0], 2, 512, 32), dtype=torch.float32)


@torch.library.register_fake("m2h::rir_convolve")
def _(mono, rir, rir_len):
    B, K, N = mono.shape
    return mono.new_empty((B, K, 2, N))


@torch.library.register_fake("m2h::stft_waves")
def _(waves, n_gt, sim_dtypes, want_phase, raw_mag):
    B, K, _, N = waves.shape
    dt = torch.float16 if sim_dtypes else torch.float32
    mix = waves.new_empty((B, 512, _T(N), 2), dtype=dt)
    phase = waves.new_empty((B, 512, _T(N), 2), dtype=dt) if want_phase else waves.new_empty(0, dtype=torch.float32)
    gt = waves.new_empty((B, n_gt, 512, _T(N), 4 if want_phase else 2), dtype=dt) if n_gt else waves.new_empty(0, dtype=torch.float32)
    return mix, phase, gt


@torch.library.register_fake("m2h::unet_forward")
def _(obs, target_class, enc_packed, dec_packed):
    return obs.new_empty((obs.shape[0], 2, 512, 32), dtype=torch.float32)


@torch.library.register_fake("m2h::bin2mono_forward")
def _(masks, obs, enc_packed, dec_packed):
    return obs.new_empty((obs.shape[0], 1, 512, 32), dtype=torch.float32)


@torch.library.register_fake("m2h::acoustic_mem")
def _(pred_mono, prev, conv1_w, conv2_w, bn_w, bn_b, bn_mean, bn_var, bn_eps):
    return pred_mono.new_empty((pred_mono.shape[0], 512, 32, 1), dtype=torch.float32)


@torch.library.register_fake("m2h::render_mixture")
def _(mono, rir, rir_len, n_gt, sim_dtypes, want_phase):
    B, K, N = mono.shape
    waves = mono.new_empty((B, K, 2, N))
    dt = torch.float16 if sim_dtypes else torch.float32
    mix = mono.new_empty((B, 512, _T(N), 2), dtype=dt)
    phase = mono.new_empty((B, 512, _T(N), 2), dtype=dt) if want_phase else mono.new_empty(0, dtype=torch.float32)
    gt = mono.new_empty((B, n_gt, 512, _T(N), 4 if want_phase else 2), dtype=dt) if n_gt else mono.new_empty(0, dtype=torch.float32)
    return mix, phase, gt, waves


@torch.library.register_fake("m2h::istft")
def _(mag, phase, length):
    return mag.new_empty((mag.shape[0], mag.shape[3], length))


@torch.library.register_fake("m2h::stft_l2_distance")
def _(mixed_audio, masks, gt_bin_comps, pred_mono, gt_mono_comps):
    B = mixed_audio.shape[0]
    return mixed_audio.new_empty((B, 1)), mixed_audio.new_empty((B, 1))
