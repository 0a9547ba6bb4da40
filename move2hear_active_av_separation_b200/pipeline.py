"""Host-buffer entry of the whole hot path: pinned host arrays in -> ratio masks in pinned host memory out.

This is the batched, in-process call SURVEY.md hard part 6 asks for next to the per-process drop-in shims: the same
arithmetic as ``get_current_mixed_bin_audio_mag_spec`` (habitat_audio/simulator_train.py:386-486, dataset variant
audio_separation/pretrain/datasets/dataset.py:162-228) followed by ``Policy.get_binSepMasks``
(audio_separation/pretrain/passive/policy.py:61-65), for a batch of mixtures whose inputs live in HOST memory.

The batch is cut into chunks that travel through ``lanes`` independent CUDA streams
(H2D copy -> RIR conv -> STFT -> U-Net -> D2H copy per chunk), so that the PCIe copies of one chunk overlap the
kernels of another: on a PCIe-attached B200 the step is bound by the host->device copy of the RIRs (256 KB per
2-source mixture), not by the kernels.  Every lane owns its own device staging buffers and its own render workspace
(the conv kernel's spectrum scratch is per launch).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from .render import AudioRenderer


class _Lane:
    def __init__(self, device, chunk: int, K: int, N: int, Lmax: int):
        self.stream = torch.cuda.Stream(device)
        self.rend = AudioRenderer(device)
        self.mono = torch.empty((chunk, K, N), dtype=torch.int16, device=device)
        self.rir = torch.empty((chunk, K, 2, Lmax), dtype=torch.float32, device=device)
        self.len = torch.empty((chunk, K), dtype=torch.int32, device=device)
        self.tc = torch.empty((chunk, 1), dtype=torch.int64, device=device)
        self.waves = torch.empty((chunk, K, 2, N), dtype=torch.int16, device=device)
        # the separator's first-layer operand, written by the STFT kernel (zero border established here, once)
        self.x0 = torch.zeros((chunk, 18, 18, 128), dtype=torch.bfloat16, device=device) if 1 + (N - 1) // 512 == 32 else None
        self.done = torch.cuda.Event()


class SeparationPipeline:
    """``pipe(mono, rir, rir_len, target_class, out)``: render + separate a host-resident batch.

    Args:
        enc, dec: ``PassiveSepEncCNN`` / ``PassiveSepDecCNN`` (binaural separator, ``.eval()``) on ``device``.
        chunk: mixtures per chunk (148 mixtures x K = 2 sources = one item per CTA of the convolution's 296-CTA grid; small
            chunks shorten the un-overlapped first copy-in and last copy-out).
        lanes: number of chunks in flight (3 = one copying in, one computing, one copying out).
    """

    def __init__(self, enc, dec, device, K: int, N: int, Lmax: int, chunk: int = 148, lanes: int = 3, stream_calls: bool = False):
        """``stream_calls``: successive calls overlap - a lane starts the next call's chunk as soon as its own previous chunk is
        done instead of waiting for the caller's stream (which waits for the whole previous call): the first copy-in of call
        i + 1 then runs under the last kernels and copy-out of call i.  Safe when the inputs are host buffers that are ready
        when the call is made (they are: pinned host memory) and the caller reads ``out`` only after synchronising its stream."""
        self.stream_calls = stream_calls
        if not torch.cuda.is_available():
            raise _lib.M2HError("no CUDA device: the B200 hot path has no CPU fallback")
        self.device = torch.device(device)
        self.enc, self.dec = enc, dec
        self.K, self.N, self.Lmax, self.chunk = K, N, Lmax, chunk
        with torch.cuda.device(self.device):
            self.lanes: List[_Lane] = [_Lane(self.device, chunk, K, N, Lmax) for _ in range(lanes)]
            # weight packing happens once, on the current stream; make it visible to every lane
            enc._pack(self.device)
            dec._pack(self.device)
            torch.cuda.current_stream(self.device).synchronize()

    @staticmethod
    def _pinned(t: torch.Tensor, name: str, dtype) -> torch.Tensor:
        if t.is_cuda or t.dtype != dtype or not t.is_contiguous():
            raise _lib.M2HError(f"{name} must be a contiguous host tensor of {dtype}")
        if not t.is_pinned():
            raise _lib.M2HError(f"{name} must live in pinned host memory (torch.Tensor.pin_memory()) for asynchronous copies")
        return t

    @torch.no_grad()
    def __call__(self, mono: torch.Tensor, rir: torch.Tensor, rir_len: torch.Tensor, target_class: torch.Tensor,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int16 ``[B,K,N]``, fp32 ``[B,K,2,Lmax]``, int32 ``[B,K]``, int64 ``[B,1]`` (pinned host) ->
        fp32 ``[B,2,512,32]`` pinned host masks (``out.permute(0,2,3,1)`` is the reference's ``[B,512,32,2]`` view).
        Returns after enqueueing; the caller's current stream waits for every lane."""
        mono = self._pinned(mono, "mono", torch.int16)
        rir = self._pinned(rir, "rir", torch.float32)
        rir_len = self._pinned(rir_len, "rir_len", torch.int32)
        target_class = self._pinned(target_class, "target_class", torch.int64)
        B = mono.shape[0]
        if tuple(mono.shape[1:]) != (self.K, self.N) or tuple(rir.shape) != (B, self.K, 2, self.Lmax):
            raise _lib.M2HError(f"shape mismatch: mono {tuple(mono.shape)} rir {tuple(rir.shape)}")
        if out is None:
            out = torch.empty((B, 2, 512, 32), dtype=torch.float32).pin_memory()
        out = self._pinned(out, "out", torch.float32)
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for i, s in enumerate(range(0, B, self.chunk)):
            e = min(B, s + self.chunk)
            n = e - s
            lane = self.lanes[i % len(self.lanes)]
            with torch.cuda.stream(lane.stream):
                if not self.stream_calls:
                    lane.stream.wait_event(start)
                lane.mono[:n].copy_(mono[s:e], non_blocking=True)
                lane.rir[:n].copy_(rir[s:e], non_blocking=True)
                lane.len[:n].copy_(rir_len[s:e], non_blocking=True)
                lane.tc[:n].copy_(target_class[s:e], non_blocking=True)
                waves = lane.rend.rir_convolve(lane.mono[:n], lane.rir[:n], lane.len[:n], out=lane.waves[:n])
                if lane.x0 is not None:       # STFT -> bf16 space-to-depth operand: no fp32 observation, no repacking kernel
                    mix = lane.rend.stft_waves_x0(waves, out=lane.x0[:n])
                else:
                    mix, _, _ = lane.rend.stft_waves(waves)
                mask = self.dec(*self.enc({"mixed_bin_audio_mag": mix, "target_class": lane.tc[:n]}))
                out[s:e].copy_(mask.permute(0, 3, 1, 2), non_blocking=True)
                lane.done.record(lane.stream)
        for lane in self.lanes:
            cur.wait_event(lane.done)
        return out
