"""CUDA-graph replay of the rollout step (graphs.GraphedSeparation) against the eager path."""
import numpy as np
import pytest
import torch

from oracle import render as R
from oracle import unet as U

pytestmark = pytest.mark.gpu


def test_graphed_rollout_step_equals_eager():
    from move2hear_active_av_separation_b200.graphs import GraphedSeparation
    from move2hear_active_av_separation_b200.render import AudioRenderer
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
    dev = torch.device("cuda", 0)
    enc_sd, dec_sd = U.init_state(False, seed=3)
    enc, dec = PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval()
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)
    B, K = 14, 2                                                  # BASELINE config 4: 14 environments per rank
    rend = AudioRenderer(dev)
    g = GraphedSeparation(enc, dec, dev, B, K, renderer=rend)
    for seed in (0, 1, 2):
        monos = torch.from_numpy(R.synth_monos(B, K, seed=seed))
        rirs, lens = R.synth_rirs(B, K, L=16000, seed=10 + seed, ragged=True)
        rirs, lens = torch.from_numpy(rirs), torch.from_numpy(lens)
        tc = torch.from_numpy(np.random.default_rng(seed).integers(0, 102, (B, 1)))
        mask = g(monos, rirs, lens, tc).clone()
        obs = g.obs.clone()
        waves = rend.rir_convolve(monos.to(dev), rirs.to(dev), lens.to(dev))
        mix, _, _ = rend.stft_waves(waves)
        with torch.no_grad():
            want = dec(*enc({"mixed_bin_audio_mag": mix, "target_class": tc.to(dev)}))
        assert torch.equal(obs, mix)
        assert mask.shape == want.shape and mask.stride() == want.stride() and torch.equal(mask, want)
    with pytest.raises(Exception):
        GraphedSeparation(enc.train(), dec, dev, B, K, renderer=rend)
