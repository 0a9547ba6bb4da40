"""Boundary wrappers that fix the module tree / state_dict names and the two call signatures the trainers use.

Mirrors ``audio_separation/pretrain/passive/policy.py:7-97`` (``PassiveSepEnc``, ``PassiveSepDec``, ``Policy``,
``Move2HearPassiveWoMemoryPolicy``) and the separator part of ``audio_separation/rl/ppo/policy.py:121-157, 183-193``
(``get_binSepMasks`` / ``convert_bin2mono``).  The ``torchsummary`` printout of the reference constructor is dropped.
"""
import torch.nn as nn

from .separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN


class PassiveSepEnc(nn.Module):
    r"""Network which encodes separated bin or mono outputs"""

    def __init__(self, observation_space=None, world_rank=0, convert_bin2mono=False):
        super().__init__()
        if observation_space is not None:
            assert "mixed_bin_audio_mag" in observation_space.spaces
        self.passive_sep_encoder = PassiveSepEncCNN(convert_bin2mono=convert_bin2mono)

    def forward(self, observations, mixed_audio=None):
        return self.passive_sep_encoder(observations, mixed_audio=mixed_audio)


class PassiveSepDec(nn.Module):
    r"""Network which decodes separated bin or mono outputs feature embeddings"""

    def __init__(self, convert_bin2mono=False):
        super().__init__()
        self.passive_sep_decoder = PassiveSepDecCNN(convert_bin2mono=convert_bin2mono)

    def forward(self, bottleneck_feats, lst_skip_feats):
        return self.passive_sep_decoder(bottleneck_feats, lst_skip_feats)


class Policy(nn.Module):
    r"""Network for the passive separation in Move2Hear pretraining"""

    def __init__(self, binSep_enc, binSep_dec, bin2mono_enc, bin2mono_dec):
        super().__init__()
        self.binSep_enc = binSep_enc
        self.binSep_dec = binSep_dec
        self.bin2mono_enc = bin2mono_enc
        self.bin2mono_dec = bin2mono_dec

    def forward(self):
        raise NotImplementedError

    def get_binSepMasks(self, observations):
        bottleneck_feats, lst_skip_feats = self.binSep_enc(observations)
        return self.binSep_dec(bottleneck_feats, lst_skip_feats)

    def convert_bin2mono(self, pred_binSepMasks, mixed_audio=None):
        bottleneck_feats, lst_skip_feats = self.bin2mono_enc(pred_binSepMasks, mixed_audio=mixed_audio)
        return self.bin2mono_dec(bottleneck_feats, lst_skip_feats)


class Move2HearPassiveWoMemoryPolicy(Policy):
    def __init__(self, observation_space=None):
        super().__init__(PassiveSepEnc(observation_space=observation_space), PassiveSepDec(),
                         PassiveSepEnc(observation_space=observation_space, convert_bin2mono=True),
                         PassiveSepDec(convert_bin2mono=True))
