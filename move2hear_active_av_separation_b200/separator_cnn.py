"""Drop-in ``PassiveSepEncCNN`` / ``PassiveSepDecCNN`` backed by the sm_100a tcgen05 kernels.

Mirrors ``audio_separation/rl/models/separator_cnn.py`` (ctor :36-54 / :119-137, forward :70-108 / :153-170):
same constructor arguments, the same module tree (``cnn.{i}.0`` conv / deconv, ``cnn.{i}.1`` BatchNorm2d,
``cnn.5.0`` 1x1 head) so that ``state_dict`` keys, shapes, ``.parameters()``, ``.train()/.eval()``, ``.to()`` and
``requires_grad_`` behave exactly like the reference's modules and its checkpoints load unchanged; the same forward
signatures and return shapes.  The torch layers only *hold* the parameters - the arithmetic runs in
``libm2h_b200.so`` (``csrc/unet_gemm.cuh``) through the C ABI of ``include/m2h.h``.

``.eval()`` modules run the inference-statistics forward (RL rollout / update_sep / evaluation call sites,
ppo_trainer.py:300-307, 359-364, ppo.py:197-204) with BatchNorm folded into the GEMM epilogues.  ``.train()`` modules run
the batch-statistics forward (running statistics updated in place like ``nn.BatchNorm2d``) and are differentiable through
``torch.autograd`` (custom Functions whose backward calls the CUDA data-/weight-gradient kernels) - the pretraining
call sites passive_trainer.py:226-236.  There is no torch fallback in either mode.
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


class _EncGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 5) for n in ("conv_w", "bn_w", "bn_b")]


class _DecGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 5) for n in ("deconv_w", "bn_w", "bn_b")] + \
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


class PackedMixture:
    """The mixture spectrogram in the layout the binaural separator's first layer reads: ``x0`` = bf16 ``[B,18,18,128]``
    (space-to-depth, haloed NHWC of ``log1p |STFT|``), written directly by the STFT kernel (``AudioRenderer.stft_waves_x0``).
    Passed as ``observations["mixed_bin_audio_mag"]`` to ``PassiveSepEncCNN`` (eval mode) instead of the fp32
    ``[B,512,32,2]`` tensor: the observation skips its fp32 round trip through HBM and the repacking kernel.  ``mag`` holds the
    fp32 observation when the caller also asked for it.  On this route the encoder hands the int64 ``target_class`` to its first
    kernel as is and returns the bottleneck features as the bf16 tensor the decoder reads (the fp32 copy of the reference's
    return value is one ``.float()`` away): no launch sits between the STFT kernel and the first convolution."""
    __slots__ = ("x0", "mag")

    def __init__(self, x0: torch.Tensor, mag: Optional[torch.Tensor] = None):
        if x0.dtype != torch.bfloat16 or x0.dim() != 4 or tuple(x0.shape[1:]) != (18, 18, 128) or not x0.is_contiguous() or not x0.is_cuda:
            raise _lib.M2HError(f"PackedMixture: expected a contiguous CUDA bf16 [B,18,18,128] tensor, got {x0.dtype} {tuple(x0.shape)}")
        self.x0, self.mag = x0, mag

    @property
    def device(self):
        return self.x0.device

    @property
    def shape(self):
        return torch.Size((self.x0.shape[0], 512, 32, 2))

    def size(self, dim: int):
        return self.shape[dim]


class _Lease:
    """Held by the tensors a forward pass returns; its death tells the pool that their workspace may be reused."""
    __slots__ = ("__weakref__",)


class _WorkspacePool:
    """Per-module cache of activation workspaces.

    A workspace (0.29 MB per sample) is reused by the next forward pass once every tensor returned from the previous pass
    that used it has been released; otherwise a new one is allocated, so results keep the reference's value semantics.
    Reuse matters because a workspace's zero borders are written once, not per call (43 us per 1024 samples).
    """

    def __init__(self, max_items: int = 8):
        self.items = []
        self.max_items = max_items

    def acquire(self, B: int, device):
        """A workspace is only handed out again on the CUDA stream that used it last (stream order = safe reuse: the
        previous pass's kernels may still be running when its Python results are released)."""
        import weakref
        lease = _Lease()
        stream = torch.cuda.current_stream(device).cuda_stream
        for ent in self.items:
            ws, ref = ent
            if ws.B == B and ws.buf.device == device and ws.stream == stream and ref() is None:
                ent[1] = weakref.ref(lease)
                return ws, lease
        if len(self.items) >= self.max_items:
            self.items = [e for e in self.items if e[1]() is not None][-(self.max_items - 1):]
        ws = _Workspace(B, device)
        ws.stream = stream
        self.items.append([ws, weakref.ref(lease)])
        return ws, lease


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
        # the kernels only ever write interiors: the zero borders (= the convolutions' padding) are established once
        self.clear_halos()

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


class _TrainWorkspace:
    """Activations, raw conv outputs, gradients and packed weights of one encoder/decoder pair (train mode)."""

    # slot -> (H = W, channels) of the haloed [B, H+2, W+2, C] maps used from Python
    CAT = {1: (16, 128), 2: (8, 256), 3: (4, 512), 4: (2, 1024)}

    def __init__(self, B: int, device):
        lib = _lib.load()
        self.B = B
        self.nbytes = lib.m2h_unet_train_workspace_bytes(B)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        offs = (ctypes.c_size_t * 32)()
        _lib.check(lib.m2h_unet_train_workspace_layout(B, offs), "m2h_unet_train_workspace_layout")
        self.offsets = list(offs)
        # bumped by every train-mode forward that writes this workspace; backward refuses to run on overwritten activations
        self.generation = 0
        # set by prepack_train_weights, consumed (and cleared) by the next train-mode forward of the encoder / the decoder
        self.prepacked_enc = self.prepacked_dec = False
        with torch.cuda.device(device):
            rc = lib.m2h_unet_train_workspace_init(B, self.buf.data_ptr(), self.nbytes, torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_train_workspace_init")

    def _map(self, slot: int, hw: int, c: int):
        n = self.B * (hw + 2) * (hw + 2) * c
        t = self.buf[self.offsets[slot]: self.offsets[slot] + 2 * n].view(torch.bfloat16).view(self.B, hw + 2, hw + 2, c)
        return t[:, 1:-1, 1:-1, :]

    def _vec(self, slot: int):
        return self.buf[self.offsets[slot]: self.offsets[slot] + 2 * self.B * 512].view(torch.bfloat16).view(self.B, 512)

    def view(self, idx: int, shape):      # same contract as _Workspace.view (slots 1-5)
        return self._vec(idx) if len(shape) == 2 else self._map(idx, shape[1], shape[3])

    def skips(self, base: int = 0) -> List[torch.Tensor]:
        """NCHW views of the skip features (base 0) or of their gradients (base 14: the dcat slots), deepest first."""
        out = []
        for lvl in (4, 3, 2, 1):
            hw, c2 = self.CAT[lvl]
            out.append(self._map(base + lvl, hw, c2)[..., c2 // 2:].permute(0, 3, 1, 2))
        return out


_WEIGHTS_PACKED = 2       # M2H_UNET_WEIGHTS_PACKED (include/m2h.h)


def prepack_train_weights(enc, dec, B: int, device) -> None:
    """Write the bf16 GEMM operands of an encoder / decoder pair into their train workspace on the CURRENT stream, ahead of the
    pair's next train-mode forward (which then skips its own packing launches).  A training loop calls this on a side stream
    while the step's audio is still being rendered and makes the forward's stream wait for it (``pretrain.GraphedTrainStep``).
    Valid for exactly one forward of each module; the parameters must not change in between."""
    lib = _lib.load()
    device = torch.device(device)
    for m in (enc, dec):
        m._check_params_on(device)
        m._check_train_dtypes()
    ws = enc._train_workspace(B, device)
    pe, keep_e = enc._param_struct()
    pd, keep_d = dec._param_struct()
    with torch.cuda.device(device):
        rc = lib.m2h_unet_pack_train_weights(ctypes.byref(pe), ctypes.byref(pd), int(enc._convert_bin2mono), B, ws.buf.data_ptr(), ws.nbytes,
                                             torch.cuda.current_stream(device).cuda_stream)
    _lib.check(rc, "m2h_unet_pack_train_weights")
    ws.prepacked_enc = ws.prepacked_dec = True


def _check_generation(ctx, ws):
    """The saved activations of a train-mode forward live in the module's workspace, not in autograd's graph: a second
    train-mode forward of the same module (gradient accumulation, a logging forward under no_grad) would overwrite them
    and backward would silently produce wrong gradients - refuse instead (the reference's nn.Module has value semantics)."""
    if ctx.generation != ws.generation:
        raise _lib.M2HError("backward through a train-mode separator forward whose activations were overwritten by a later "
                            "train-mode forward of the same module: call backward() before the next forward (or use a "
                            "second module / eval mode for the extra forward)")


def _check_device(device, **tensors):
    """Raw pointers go to the kernels: a tensor on the host or on another GPU would fault the whole CUDA context."""
    for name, t in tensors.items():
        if t is not None and t.device != device:
            raise _lib.M2HError(f"{name} is on {t.device} but the separator input is on {device}: move the module and all of "
                                "its inputs to the same CUDA device")


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * 5)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _EncTrainFn(torch.autograd.Function):
    """Encoder forward with batch statistics; backward = BatchNorm backward + weight gradients + skip data gradients."""

    @staticmethod
    def forward(ctx, module, obs, cls, masks, *params):
        lib = _lib.load()
        device = obs.device
        B = obs.size(0)
        ws = module._train_workspace(B, device)
        ws.generation += 1
        p, keep = module._param_struct()
        nbt = _ptr_array([module.cnn[i][1].num_batches_tracked for i in range(5)])
        bott = torch.empty((B, 512), dtype=torch.float32, device=device)
        strides = (ctypes.c_int64 * 4)(*masks.stride()) if masks is not None else None
        bn = module.cnn[0][1]
        mode = int(module._convert_bin2mono) | (_WEIGHTS_PACKED if ws.prepacked_enc else 0)
        ws.prepacked_enc = False
        with torch.cuda.device(device):
            rc = lib.m2h_unet_encoder_forward_train(
                ctypes.byref(p), nbt, mode, float(bn.eps), float(bn.momentum), obs.data_ptr(),
                cls.data_ptr() if cls is not None else None, masks.data_ptr() if masks is not None else None, strides, B,
                ws.buf.data_ptr(), ws.nbytes, bott.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_encoder_forward_train")
        ctx.module, ctx.ws, ctx.cls, ctx.B, ctx.device = module, ws, cls, B, device
        ctx.generation = ws.generation
        skips = ws.skips()
        bott._m2h_ws = ws
        return (bott, *skips)

    @staticmethod
    def backward(ctx, d_bott, *d_skips):
        lib = _lib.load()
        module, ws, B, device = ctx.module, ctx.ws, ctx.B, ctx.device
        _check_generation(ctx, ws)
        # gradients of the outputs must sit in the workspace slots the kernels read (they already do when they come from _DecTrainFn)
        own = ws.skips(base=14)
        for g, slot in zip(d_skips, own):
            if g is None:
                slot.zero_()
            elif g.data_ptr() != slot.data_ptr() or g.stride() != slot.stride():
                slot.copy_(g)
        dbott = ws._vec(19)
        if d_bott is None:
            dbott.zero_()
        else:
            dbott.copy_(d_bott)
        p, keep = module._param_struct()
        grads, gstruct = module._grad_struct(device)
        with torch.cuda.device(device):
            rc = lib.m2h_unet_encoder_backward(ctypes.byref(p), int(module._convert_bin2mono),
                                               ctx.cls.data_ptr() if ctx.cls is not None else None, B, ws.buf.data_ptr(), ws.nbytes,
                                               ctypes.byref(gstruct), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_encoder_backward")
        return (None, None, None, None, *[g if q.requires_grad else None for g, q in zip(grads, module._train_params())])


class _DecTrainFn(torch.autograd.Function):
    """Decoder forward with batch statistics (-> fp32 [B, C, 512, 32]); backward leaves d(bottleneck) / d(skips) in the workspace."""

    @staticmethod
    def forward(ctx, module, ws, bott, s0, s1, s2, s3, *params):
        lib = _lib.load()
        device = bott.device
        B = bott.size(0)
        p, keep = module._param_struct()
        nbt = _ptr_array([module.cnn[i][1].num_batches_tracked for i in range(5)])
        C = module._n_out_audio // SLICE_FACTOR
        out = torch.empty((B, C, 512, 32), dtype=torch.float32, device=device)
        bn = module.cnn[0][1]
        mode = int(module._convert_bin2mono) | (_WEIGHTS_PACKED if ws.prepacked_dec else 0)
        ws.prepacked_dec = False
        with torch.cuda.device(device):
            rc = lib.m2h_unet_decoder_forward_train(ctypes.byref(p), nbt, mode, float(bn.eps), float(bn.momentum),
                                                    B, ws.buf.data_ptr(), ws.nbytes, out.data_ptr(),
                                                    torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_decoder_forward_train")
        ctx.module, ctx.ws, ctx.B, ctx.device = module, ws, B, device
        ctx.generation = ws.generation
        return out

    @staticmethod
    def backward(ctx, dmask):
        lib = _lib.load()
        module, ws, B, device = ctx.module, ctx.ws, ctx.B, ctx.device
        _check_generation(ctx, ws)
        dmask = _f32(dmask)
        p, keep = module._param_struct()
        grads, gstruct = module._grad_struct(device)
        with torch.cuda.device(device):
            rc = lib.m2h_unet_decoder_backward(ctypes.byref(p), int(module._convert_bin2mono), dmask.data_ptr(), B, ws.buf.data_ptr(),
                                               ws.nbytes, ctypes.byref(gstruct), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_unet_decoder_backward")
        d_bott = ws._vec(19).float()
        d_skips = ws.skips(base=14)
        return (None, None, d_bott, *d_skips, *[g if q.requires_grad else None for g, q in zip(grads, module._train_params())])


class _L1MaskLossFn(torch.autograd.Function):
    """``F.l1_loss(mask * (exp(obs) - 1), gt)`` (use_mix) or ``F.l1_loss(mask, gt)`` in one kernel (passive_trainer.py:269-275)."""

    @staticmethod
    def forward(ctx, mask, obs, gt, use_mix):
        lib = _lib.load()
        device = mask.device
        B, C = mask.size(0), mask.size(3)
        _check_device(device, gt=gt, mixed_audio=obs if use_mix else None)
        nchw = mask.permute(0, 3, 1, 2)
        if not nchw.is_contiguous() or nchw.dtype != torch.float32:
            nchw = nchw.float().contiguous()
        gt = _f32(gt)
        obs_c = _f32(obs) if use_mix else None
        numel = B * C * 512 * 32
        loss_sum = torch.empty((), dtype=torch.float64, device=device)
        dmask = torch.empty((B, C, 512, 32), dtype=torch.float32, device=device) if mask.requires_grad else None
        with torch.cuda.device(device):
            rc = lib.m2h_l1_mask_loss(nchw.data_ptr(), obs_c.data_ptr() if use_mix else None, gt.data_ptr(), B, C, int(use_mix),
                                      1.0 / numel, loss_sum.data_ptr(), dmask.data_ptr() if dmask is not None else None,
                                      torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "m2h_l1_mask_loss")
        ctx.dmask = dmask
        return (loss_sum / numel).float()

    @staticmethod
    def backward(ctx, grad_out):
        return (ctx.dmask * grad_out).permute(0, 2, 3, 1), None, None, None


def l1_mask_loss(pred_masks: torch.Tensor, gt: torch.Tensor, mixed_audio: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Mean-reduced L1 loss on CUDA: ``|pred_masks * (exp(mixed_audio) - 1) - gt|`` when ``mixed_audio`` is given, else ``|pred_masks - gt|``."""
    if not pred_masks.is_cuda:
        raise _lib.M2HError("l1_mask_loss needs CUDA tensors (there is no CPU path)")
    return _L1MaskLossFn.apply(pred_masks, mixed_audio, gt, mixed_audio is not None)


class _PackedModule(nn.Module):
    """Common parameter-packing cache: re-pack when any parameter/buffer changed (load_state_dict, optimiser step)."""

    def _signature(self):
        sig = []
        for t in list(self.parameters()) + list(self.buffers()):
            sig.append((t.data_ptr(), t._version, t.device.index, t.dtype))
        return tuple(sig)

    def _require_cuda(self, device):
        if device.type != "cuda":
            raise _lib.M2HError("the B200 U-Net has no CPU path: move the module and its inputs to a CUDA device")

    # ---- train mode -----------------------------------------------------------------------------------------------
    def _train_workspace(self, B, device):
        ws = getattr(self, "_tws", None)
        if ws is None or ws.B != B or ws.buf.device != device:
            ws = _TrainWorkspace(B, device)
            self._tws = ws
        return ws

    def _alloc_grads(self):
        """Gradient buffers of one backward: fresh tensors, or - under data-parallel training - fresh views of the
        ``distributed.GradientAllReducer`` arena, so that the all-reduce runs in place on what the kernels wrote."""
        arena = getattr(self, "_grad_arena", None)
        ps = self._train_params()
        if arena is not None and all(q.requires_grad and q.dtype == torch.float32 for q in ps):
            # a parameter that still holds a gradient (accumulation over several backward passes without zero_grad) keeps it:
            # the kernels must not overwrite the arena slot p.grad lives in - that backward gets a separate buffer and autograd
            # adds it into the slot
            return [arena.grad_view(q) if q.grad is None else torch.empty_like(q, memory_format=torch.contiguous_format) for q in ps]
        return [torch.empty_like(q, memory_format=torch.contiguous_format) for q in ps]

    def _train_params(self):
        """Parameters in the order the autograd Functions return their gradients: conv weights, BN weights, BN biases (+ head)."""
        ps = [self.cnn[i][0].weight for i in range(5)] + [self.cnn[i][1].weight for i in range(5)] + [self.cnn[i][1].bias for i in range(5)]
        if len(self.cnn) > 5:
            ps += [self.cnn[5][0].weight, self.cnn[5][0].bias]
        return ps

    def _check_params_on(self, device):
        for name, t in list(self.named_parameters()) + list(self.named_buffers()):
            if t.device != device:
                raise _lib.M2HError(f"parameter / buffer {name} is on {t.device} but the input is on {device}: call "
                                    f".to({str(device)!r}) on the module first")

    def _check_train_dtypes(self):
        for t in list(self.parameters()) + [b for b in self.buffers() if b.dtype.is_floating_point]:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.M2HError("train mode needs contiguous fp32 parameters / buffers (running statistics are updated in place)")


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

    def _param_struct(self):
        keep, p = [], _EncParams()
        for i in range(5):
            conv, bn = self.cnn[i][0], self.cnn[i][1]
            for name, t in (("conv_w", conv.weight), ("bn_w", bn.weight), ("bn_b", bn.bias),
                            ("bn_mean", bn.running_mean), ("bn_var", bn.running_var)):
                keep.append(t)
                getattr(p, name)[i] = t.data_ptr()
        return p, keep

    def _grad_struct(self, device):
        grads = self._alloc_grads()
        g = _EncGrads()
        for i in range(5):
            g.conv_w[i], g.bn_w[i], g.bn_b[i] = grads[i].data_ptr(), grads[5 + i].data_ptr(), grads[10 + i].data_ptr()
        return grads, g

    def _forward_train(self, obs, observations, masks, B, device):
        self._check_train_dtypes()
        cls = None
        if self._convert_bin2mono:
            masks = masks.detach()                     # the converter consumes detached masks (passive_trainer.py:229-230)
            if masks.dtype != torch.float32:
                masks = masks.float()
        else:
            cls = (observations["target_class"].detach().reshape(B).float() + 1).contiguous()
        outs = _EncTrainFn.apply(self, obs, cls, masks, *self._train_params())
        outs[0]._m2h_ws = self._tws                    # lets the partner decoder find the activations (and leave its gradients) in place
        return outs[0], list(outs[1:])

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
        self._require_cuda(device)
        self._check_params_on(device)
        _check_device(device, masks=masks, target_class=None if self._convert_bin2mono else observations["target_class"])
        packed_obs = obs if isinstance(obs, PackedMixture) else None
        if packed_obs is not None:
            if self._convert_bin2mono or self.training:
                if obs.mag is None:
                    raise _lib.M2HError("a PackedMixture without .mag feeds the binaural separator in eval mode only")
                obs, packed_obs = obs.mag, None
        if packed_obs is None:
            obs = _f32(obs)
        B = obs.size(0)
        if tuple(obs.shape[1:]) != (512, 32, 2):
            raise _lib.M2HError(f"expected a [B,512,32,2] spectrogram, got {tuple(obs.shape)}")
        if self.training:
            return self._forward_train(obs, observations, masks, B, device)
        packed = self._pack(device)
        if getattr(self, "_ws_pool", None) is None:
            self._ws_pool = _WorkspacePool()
        ws, lease = self._ws_pool.acquire(B, device)
        cls_ptr = mask_ptr = strides = None
        keep = []
        if packed_obs is not None:
            # fused hand-over: raw int64 classes go to the kernel (no conversion launch between the STFT and the first layer) and
            # the bottleneck features are returned as the workspace's bf16 copy (no fp32 conversion launch)
            cls = observations["target_class"].detach().reshape(B)
            if cls.dtype != torch.int64 or not cls.is_contiguous():
                cls = cls.to(torch.int64).contiguous()
            keep.append(cls)
            bott = ws.view(5, (B, 512))
            with torch.cuda.device(device):
                rc = lib.m2h_unet_encoder_forward_x0(packed.data_ptr(), packed_obs.x0.data_ptr(), cls.data_ptr(), B, ws.buf.data_ptr(), ws.nbytes,
                                                     None, torch.cuda.current_stream(device).cuda_stream)
        else:
            bott = torch.empty((B, 512), dtype=torch.float32, device=device)
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
        bott._m2h_lease = lease
        for t in skips:
            t._m2h_lease = lease
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

    def _param_struct(self):
        keep, p = [], _DecParams()
        for i in range(5):
            conv, bn = self.cnn[i][0], self.cnn[i][1]
            for name, t in (("deconv_w", conv.weight), ("bn_w", bn.weight), ("bn_b", bn.bias),
                            ("bn_mean", bn.running_mean), ("bn_var", bn.running_var)):
                keep.append(t)
                getattr(p, name)[i] = t.data_ptr()
        p.head_w, p.head_b = self.cnn[5][0].weight.data_ptr(), self.cnn[5][0].bias.data_ptr()
        return p, keep

    def _grad_struct(self, device):
        grads = self._alloc_grads()
        g = _DecGrads()
        for i in range(5):
            g.deconv_w[i], g.bn_w[i], g.bn_b[i] = grads[i].data_ptr(), grads[5 + i].data_ptr(), grads[10 + i].data_ptr()
        g.head_w, g.head_b = grads[15].data_ptr(), grads[16].data_ptr()
        return grads, g

    def _forward_train(self, bottleneck_feats, lst_skip_feats, B, device):
        self._check_train_dtypes()
        ws = getattr(bottleneck_feats, "_m2h_ws", None)
        if not isinstance(ws, _TrainWorkspace) or ws.B != B or ws.buf.device != device:
            # features did not come straight from the partner encoder in train mode: stage them into a fresh training workspace
            ws = self._train_workspace(B, device)
            ws.generation += 1
            ws.view(5, (B, 512)).copy_(bottleneck_feats.detach().reshape(B, 512))
            for idx, feat, (hw, c) in zip((4, 3, 2, 1), lst_skip_feats, ((2, 512), (4, 256), (8, 128), (16, 64))):
                ws.view(idx, (B, hw, hw, 2 * c))[..., c:].copy_(feat.detach().permute(0, 2, 3, 1))
        out = _DecTrainFn.apply(self, ws, bottleneck_feats, *lst_skip_feats, *self._train_params())
        return out.permute(0, 2, 3, 1)

    def forward(self, bottleneck_feats, lst_skip_feats):
        lib = _lib.load()
        device = bottleneck_feats.device
        self._require_cuda(device)
        self._check_params_on(device)
        _check_device(device, **{"skip%d" % i: t for i, t in enumerate(lst_skip_feats)})
        B = bottleneck_feats.size(0)
        if self.training:
            return self._forward_train(bottleneck_feats, lst_skip_feats, B, device)
        ws: Optional[_Workspace] = getattr(bottleneck_feats, "_m2h_ws", None)
        if isinstance(ws, _TrainWorkspace):
            ws = None
        if ws is None or ws.B != B or ws.buf.device != device:
            # features did not come straight from the partner encoder: stage them (layout/dtype copies only)
            ws = _Workspace(B, device)
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
