"""CUDA-graph replay of the fixed-shape rollout step: render + separator inference for the B environments of a rank.

In the reference's RL loop every environment step renders one mixture per env worker and the trainer runs the separator on
the stacked observations (audio_separation/rl/ppo/ppo_trainer.py:296-319; BASELINE config 4: 8 ranks x 14 envs).  At
B = 14 the kernels take ~0.2 ms while ~15 launches plus the Python around them take longer, so the whole step - RIR
convolution, STFT, encoder, decoder - is captured once into a CUDA graph and replayed: one launch per step.  (CUDA streams
and graphs, not a tracing compiler: the graph contains exactly the kernels of the eager path.)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .render import AudioRenderer


class GraphedSeparation:
    """``(mono, rir, rir_len, target_class) -> mask`` for a fixed batch shape, replayed from a CUDA graph.

    Inputs are copied into static device buffers (host or device tensors accepted); the returned mask - the reference's
    ``[B, 512, 32, C]`` view - lives in graph-owned memory and is overwritten by the next call.
    """

    def __init__(self, enc, dec, device, B: int, K: int = 2, N: int = 16000, Lmax: int = 16000,
                 renderer: Optional[AudioRenderer] = None, warmup: int = 3, bin2mono_enc=None, bin2mono_dec=None,
                 post=None, acoustic_mem=None, with_stft_l2: bool = False):
        """``bin2mono_enc`` / ``bin2mono_dec`` (optional) add the second U-Net of the rollout step,
        ``pred_mono = convert_bin2mono(mask, mixed_audio=obs)`` (rl/ppo/ppo_trainer.py:304-308), to the graph;
        ``post`` (optional callable ``(obs, mask, pred_mono) -> anything``, e.g. the AcousticMem refiner of
        rl/models/memory_nets.py on static buffers) is captured after it; results in ``.pred_mono`` / ``.post_out``.

        ``acoustic_mem`` (a ``postproc.AcousticMem`` in eval mode) adds the memory refiner of the rollout step
        (``pred_monoFromMem = acoustic_mem(pred_mono, prev_pred_monoFromMem_masked)``, ppo_trainer.py:310-319) as one fused
        kernel: write the masked previous prediction into ``.prev_mem`` before the call, read ``.pred_monoFromMem`` after.
        ``with_stft_l2`` adds ``STFT_L2_distance`` (eval_metrics.py:306-366; ppo_trainer.py:407-420) on the static
        ``.gt_bin_comps`` ``[B,512,T,4]`` / ``.gt_mono_comps`` ``[B,512,T,2]`` buffers: ``.stft_l2`` = dict of device
        ``[B,1]`` tensors ``bin`` / ``mono`` (and ``monoFromMem`` with a refiner) - the whole post-mask part of the step
        replays from the same graph."""
        if not torch.cuda.is_available():
            raise _lib.M2HError("no CUDA device: the B200 path has no CPU fallback")
        mods = [m for m in (enc, dec, bin2mono_enc, bin2mono_dec) if m is not None]
        if any(m.training for m in mods):
            raise _lib.M2HError("GraphedSeparation captures the inference path: call .eval() on the separator first")
        if (bin2mono_enc is None) != (bin2mono_dec is None):
            raise _lib.M2HError("bin2mono_enc and bin2mono_dec come as a pair")
        if (acoustic_mem is not None or with_stft_l2) and bin2mono_enc is None:
            raise _lib.M2HError("the memory refiner and STFT_L2_distance consume pred_mono: pass bin2mono_enc / bin2mono_dec")
        if acoustic_mem is not None and acoustic_mem.training:
            raise _lib.M2HError("GraphedSeparation captures the inference path: call .eval() on the acoustic memory first")
        self.device = torch.device(device)
        self.enc, self.dec = enc, dec
        self.enc2, self.dec2, self.post = bin2mono_enc, bin2mono_dec, post
        self.pred_mono = None
        self.post_out = None
        self.acoustic_mem, self.with_stft_l2 = acoustic_mem, with_stft_l2
        self.pred_monoFromMem = None
        self.stft_l2 = None
        self.rend = renderer or AudioRenderer(self.device)
        dev = self.device
        self.mono = torch.zeros((B, K, N), dtype=torch.int16, device=dev)
        self.rir = torch.zeros((B, K, 2, Lmax + (Lmax & 1)), dtype=torch.float32, device=dev)
        self.rir_len = torch.zeros((B, K), dtype=torch.int32, device=dev)
        self.target_class = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        self._waves = torch.empty((B, K, 2, N), dtype=torch.int16, device=dev)
        T = 1 + (N - 1) // 512
        self.prev_mem = torch.zeros((B, 512, T, 1), dtype=torch.float32, device=dev) if acoustic_mem is not None else None
        self.gt_bin_comps = torch.zeros((B, 512, T, 4), dtype=torch.float32, device=dev) if with_stft_l2 else None
        self.gt_mono_comps = torch.zeros((B, 512, T, 2), dtype=torch.float32, device=dev) if with_stft_l2 else None
        self.obs = None
        self.mask = None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # warm-up outside the capture: weight packing, lazy module state
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.obs, self.mask, self.pred_mono, self.post_out, self.pred_monoFromMem, self.stft_l2 = self._step()

    def _step(self):
        waves = self.rend.rir_convolve(self.mono, self.rir, self.rir_len, out=self._waves)
        mix, _, _ = self.rend.stft_waves(waves)
        pred_mono = post_out = None
        with torch.no_grad():
            mask = self.dec(*self.enc({"mixed_bin_audio_mag": mix, "target_class": self.target_class}))
            if self.enc2 is not None:
                pred_mono = self.dec2(*self.enc2(mask, mixed_audio=mix))
            if self.post is not None:
                post_out = self.post(mix, mask, pred_mono)
            mem = l2 = None
            if self.acoustic_mem is not None:
                mem = self.acoustic_mem(pred_mono, self.prev_mem)
            if self.with_stft_l2:
                from .postproc import STFT_L2_distance
                b, m = STFT_L2_distance(mix, mask, self.gt_bin_comps, pred_mono, self.gt_mono_comps, to_cpu=False)
                l2 = {"bin": b, "mono": m}
                if mem is not None:
                    l2["monoFromMem"] = STFT_L2_distance(mix, mask, self.gt_bin_comps, mem, self.gt_mono_comps, to_cpu=False)[1]
        return mix, mask, pred_mono, post_out, mem, l2

    def __call__(self, mono, rir, rir_len, target_class):
        L = rir.shape[-1]
        if L > self.rir.shape[-1]:
            raise _lib.M2HError(f"RIR of {L} taps exceeds the captured Lmax = {self.rir.shape[-1]}")
        self.mono.copy_(mono, non_blocking=True)
        self.rir[..., :L].copy_(rir, non_blocking=True)          # taps beyond rir_len are never read
        self.rir_len.copy_(rir_len, non_blocking=True)
        self.target_class.copy_(target_class.reshape(self.target_class.shape), non_blocking=True)
        self.graph.replay()
        return self.mask
