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


def train_step(actor_critic, optimizer, mixed_audio, gt_bin_mag, gt_mono_mag, target_class, grad_sync=None):
    """One train-split iteration of ``PassiveTrainer.train`` (passive_trainer.py:218-245) -> ``(bin_loss, mono_loss)``."""
    obs_batch = {"mixed_bin_audio_mag": mixed_audio, "target_class": target_class}
    pred_binSepMasks = actor_critic.get_binSepMasks(obs_batch)
    pred_mono = actor_critic.convert_bin2mono(pred_binSepMasks.detach(), mixed_audio=mixed_audio)
    bin_loss, mono_loss, _ = optimize_supervised_loss(optimizer, mixed_audio, pred_binSepMasks, gt_bin_mag, pred_mono, gt_mono_mag,
                                                      split="train", grad_sync=grad_sync)
    return bin_loss, mono_loss
