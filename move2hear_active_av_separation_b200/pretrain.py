"""Pretraining step of the passive separators on the CUDA path.

Mirrors ``PassiveTrainer.optimize_supervised_loss`` and the train-split body of ``PassiveTrainer.train``
(audio_separation/pretrain/passive/passive_trainer.py:226-245, 269-286):

    pred_binSepMasks = actor_critic.get_binSepMasks(obs_batch)
    pred_mono = actor_critic.convert_bin2mono(pred_binSepMasks.detach(), mixed_audio=mixed_audio)
    bin_loss = L1(pred_binSepMasks * (exp(mixed_audio) - 1), gt_bin_mag);  mono_loss = L1(pred_mono, gt_mono_mag)
    optimizer.zero_grad(); (bin_loss + mono_loss).backward(); optimizer.step()

The reference calls ``clip_grad_norm_`` between ``zero_grad()`` and ``backward()``, where every gradient is zero /
``None`` - a no-op - so no clipping happens here either.  The L1 losses run in one fused kernel each
(``separator_cnn.l1_mask_loss``); the optimizer is whatever the caller passes (the reference: ``torch.optim.Adam(lr=5e-4, eps=1e-5)``).
"""
from __future__ import annotations

from .separator_cnn import l1_mask_loss


def optimize_supervised_loss(optimizer, mixed_audio, pred_binSepMasks, gt_bin_mag, pred_mono, gt_mono_mag, split="train",
                             grad_sync=None):
    """Same arguments / return value as the reference method; ``grad_sync`` (optional callable) runs between
    ``backward()`` and ``optimizer.step()`` - the data-parallel gradient all-reduce (``distributed.GradientAllReducer.finish``)."""
    bin_loss = l1_mask_loss(pred_binSepMasks, gt_bin_mag, mixed_audio=mixed_audio)
    mono_loss = l1_mask_loss(pred_mono, gt_mono_mag)
    if split == "train":
        optimizer.zero_grad()
        loss = bin_loss + mono_loss
        loss.backward()
        if grad_sync is not None:
            grad_sync()
        optimizer.step()
    return bin_loss, mono_loss, optimizer


def prepack_weights(actor_critic, batch_size: int, device) -> None:
    """Pack the bf16 GEMM operands of both separators for their next train-mode forward, on the current stream
    (``separator_cnn.prepack_train_weights``): issued on a side stream, this takes ~20 short launches off the head of the step."""
    from .separator_cnn import prepack_train_weights
    ac = actor_critic
    prepack_train_weights(ac.binSep_enc.passive_sep_encoder, ac.binSep_dec.passive_sep_decoder, batch_size, device)
    prepack_train_weights(ac.bin2mono_enc.passive_sep_encoder, ac.bin2mono_dec.passive_sep_decoder, batch_size, device)


def separator_param_groups(actor_critic):
    """``[separator encoder, separator decoder, converter encoder, converter decoder]`` parameters in registration order: the
    ``param_groups`` argument of ``distributed.GradientAllReducer`` (with ``start_groups_from_hooks=True``) that lets a
    group's all-reduce start as soon as its backward function has returned - the separator's under the converter's forward
    and backward, each decoder's under its encoder's backward; only the converter encoder's 28 MB stay exposed."""
    ac = actor_critic
    return [list(ac.binSep_enc.parameters()), list(ac.binSep_dec.parameters()),
            list(ac.bin2mono_enc.parameters()), list(ac.bin2mono_dec.parameters())]


def train_step(actor_critic, optimizer, mixed_audio, gt_bin_mag, gt_mono_mag, target_class, grad_sync=None, side_stream=None,
               gt_ready=None):
    """One train-split iteration of ``PassiveTrainer.train`` (passive_trainer.py:218-245) -> ``(bin_loss, mono_loss)``.

    ``side_stream`` (optional ``torch.cuda.Stream``): the binaural-to-monaural converter consumes the DETACHED masks
    (passive_trainer.py:229-230), so its forward, its loss and its whole backward are independent of the binaural separator's
    loss and backward.  With a side stream the converter branch runs there (autograd replays a node on its forward's stream),
    next to the separator's backward on the calling stream: the step is ~230 short, latency-bound kernels, and the two chains
    fill each other's gaps.  Same arithmetic, same results.

    ``gt_ready`` (optional ``torch.cuda.Event``, with ``side_stream``): the ground-truth spectrograms were produced on the side
    stream and are complete at this event - only the losses read them, so the separator's forward does not wait for them."""
    import torch
    obs_batch = {"mixed_bin_audio_mag": mixed_audio, "target_class": target_class}
    pred_binSepMasks = actor_critic.get_binSepMasks(obs_batch)
    if side_stream is None:
        pred_mono = actor_critic.convert_bin2mono(pred_binSepMasks.detach(), mixed_audio=mixed_audio)
        bin_loss, mono_loss, _ = optimize_supervised_loss(optimizer, mixed_audio, pred_binSepMasks, gt_bin_mag, pred_mono, gt_mono_mag,
                                                          split="train", grad_sync=grad_sync)
        return bin_loss, mono_loss
    main = torch.cuda.current_stream(mixed_audio.device)
    side_stream.wait_stream(main)                       # the masks (and everything before them) are ready
    with torch.cuda.stream(side_stream):
        pred_mono = actor_critic.convert_bin2mono(pred_binSepMasks.detach(), mixed_audio=mixed_audio)
        mono_loss = l1_mask_loss(pred_mono, gt_mono_mag)
    if gt_ready is not None:
        main.wait_event(gt_ready)
    bin_loss = l1_mask_loss(pred_binSepMasks, gt_bin_mag, mixed_audio=mixed_audio)
    optimizer.zero_grad()
    # (bin_loss + mono_loss).backward() of the reference = two backward passes over DISJOINT parameter sets (detached masks):
    # the separator's starts as soon as its loss exists, next to the converter's forward on the side stream, and its gradients
    # are complete - ready for the data-parallel all-reduce - while the converter's backward still runs
    bin_loss.backward()
    if grad_sync is not None and hasattr(grad_sync, "start"):
        for gi in (0, 1):                               # separator groups (separator_param_groups): reduced under the converter's backward
            grad_sync.start(gi)
    with torch.cuda.stream(side_stream):
        mono_loss.backward()
    main.wait_stream(side_stream)
    if grad_sync is not None:
        grad_sync()
    optimizer.step()
    return bin_loss, mono_loss


class GraphedTrainStep:
    """One pretraining iteration - (optionally the dataset-path render,) both separators forward with batch statistics, the two
    L1 losses, backward, (the data-parallel gradient all-reduce,) the optimizer update - captured ONCE into a CUDA graph and
    replayed per step.

    Why: at the reference's batch sizes the step is ~230 short kernels and the Python / launch path that enqueues them takes
    as long as the GPU needs to run them (B = 256 on a B200: 3.7 ms of host time for ~3 ms of GPU time per step, measured with
    ``tools/train_hostbound.py``): the step is host-bound.  Replaying a graph removes the host from the loop; the graph holds
    exactly the kernels of the eager path (CUDA streams and graphs, not a tracing compiler).

    ``optimizer`` must be capturable (``torch.optim.Adam(..., fused=True, capturable=True)``).  Inputs are copied into static
    buffers; the returned losses are static device tensors overwritten by the next call.  Drop every reference to the autograd
    graph of earlier EAGER steps first (e.g. loss tensors with a ``grad_fn``): their ``AccumulateGrad`` nodes are bound to the
    stream they ran on and invalidate the capture.  Capturing needs a few eager warm-up
    steps (lazy workspaces, optimizer state): parameters, buffers and optimizer state are restored IN PLACE afterwards, so
    the first replay is the first training step.
    """

    def __init__(self, actor_critic, optimizer, batch_size: int, device, renderer=None, n_sources: int = 2, n_samples: int = 16000,
                 rir_taps: int = 16000, grad_sync=None, warmup: int = 3, two_streams: bool = False):
        import torch
        from . import _lib
        if not torch.cuda.is_available():
            raise _lib.M2HError("no CUDA device: the B200 path has no CPU fallback")
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise _lib.M2HError("GraphedTrainStep needs a capturable optimizer: torch.optim.Adam(..., fused=True, capturable=True)")
        self.device = torch.device(device)
        dev, B = self.device, batch_size
        self.actor_critic, self.optimizer, self.rend, self.grad_sync = actor_critic, optimizer, renderer, grad_sync
        # two_streams: the converter branch of the step on a second stream (see train_step): two parallel chains in the graph
        self.side = torch.cuda.Stream(self.device) if two_streams else None
        if renderer is not None:          # render inside the graph: the dataset item's inputs are the static buffers
            self.mono = torch.zeros((B, n_sources, n_samples), dtype=torch.int16, device=dev)
            self.rir = torch.zeros((B, n_sources, 2, rir_taps + (rir_taps & 1)), dtype=torch.float32, device=dev)
            self.rir_len = torch.zeros((B, n_sources), dtype=torch.int32, device=dev)
            self.static = (self.mono, self.rir, self.rir_len)
        else:
            T = 1 + (n_samples - 1) // 512
            self.mixed_audio = torch.zeros((B, 512, T, 2), dtype=torch.float32, device=dev)
            self.gt_bin_mag = torch.zeros((B, 512, T, 2), dtype=torch.float32, device=dev)
            self.gt_mono_mag = torch.zeros((B, 512, T, 1), dtype=torch.float32, device=dev)
            self.static = (self.mixed_audio, self.gt_bin_mag, self.gt_mono_mag)
        self.target_class = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        # ---- snapshot, eager warm-up on a side stream, capture, in-place restore
        params = list(actor_critic.parameters()) + list(actor_critic.buffers())
        saved = [t.detach().clone() for t in params]
        saved_opt = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in optimizer.state.items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.bin_loss, self.mono_loss = self._step()
        with torch.no_grad():
            for t, s in zip(params, saved):
                t.copy_(s)
            for p, st in optimizer.state.items():
                old = saved_opt.get(p)
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()                 # a fresh optimizer: step 0, zero moments
        torch.cuda.synchronize(dev)

    def _step(self):
        import torch
        gt_ready = None
        if self.side is not None:
            # prologue on the side stream, next to the render / the separator's forward on the main stream: the weight packing
            # of both separators (depends on the parameters only) and the ground-truth spectrograms (read by the losses only)
            main = torch.cuda.current_stream(self.device)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                prepack_weights(self.actor_critic, self.target_class.shape[0], self.device)
                packs_ready = torch.cuda.Event()
                packs_ready.record(self.side)
        if self.rend is not None:
            mono, rir, ln = self.static                                     # dataset path (dataset.py:162-228): fp32 mixture + target GT
            if self.side is None:
                out = self.rend.render(mono, rir, ln, n_gt=1)
                mixed, gt_bin = out["mix_mag"], out["gt"][:, 0]
                gt_mono, _ = self.rend.stft_f32((mono[:, :1].float() * (1.0 / 32768)).contiguous())
            else:
                with torch.cuda.stream(self.side):                          # the clean clip's spectrogram does not need the convolution
                    gt_mono, _ = self.rend.stft_f32((mono[:, :1].float() * (1.0 / 32768)).contiguous())
                waves = self.rend.rir_convolve(mono, rir, ln)
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    _, _, gt = self.rend.stft_waves(waves, 1, gt_only=True)
                    gt_bin = gt[:, 0]
                    gt_ready = torch.cuda.Event()
                    gt_ready.record(self.side)
                mixed, _, _ = self.rend.stft_waves(waves, 0)
        else:
            mixed, gt_bin, gt_mono = self.static
        if self.side is not None:
            main.wait_event(packs_ready)
        bin_loss, mono_loss = train_step(self.actor_critic, self.optimizer, mixed, gt_bin, gt_mono, self.target_class, grad_sync=self.grad_sync,
                                         side_stream=self.side, gt_ready=gt_ready)
        return bin_loss.detach(), mono_loss.detach()      # no reference to the autograd graph survives the step

    def __call__(self, *inputs, target_class=None):
        """``(mono, rir, rir_len)`` with a renderer, else ``(mixed_audio, gt_bin_mag, gt_mono_mag)``; plus ``target_class``."""
        for dst, src in zip(self.static, inputs):
            if dst.shape != src.shape:
                if dst is getattr(self, "rir", None) and src.shape[:3] == dst.shape[:3] and src.shape[3] <= dst.shape[3]:
                    dst[..., :src.shape[3]].copy_(src, non_blocking=True)       # taps beyond rir_len are never read
                    continue
                from . import _lib
                raise _lib.M2HError(f"GraphedTrainStep was captured for {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)
        if target_class is not None:
            self.target_class.copy_(target_class.reshape(self.target_class.shape), non_blocking=True)
        self.graph.replay()
        return self.bin_loss, self.mono_loss
