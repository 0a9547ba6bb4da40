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


def test_graphed_rollout_step_with_bin2mono_and_memory_refiner():
    """Both U-Nets of the rollout step (rl/ppo/ppo_trainer.py:298-319) plus a torch post-module in one graph."""
    from move2hear_active_av_separation_b200.graphs import GraphedSeparation
    from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    pol = Move2HearPassiveWoMemoryPolicy().to(dev).eval()
    refiner = torch.nn.Sequential(torch.nn.Conv2d(32, 32, 3, padding=1, bias=False), torch.nn.ReLU(),
                                  torch.nn.Conv2d(32, 16, 3, padding=1, bias=False)).to(dev).eval()     # AcousticMem's layers (use_ddppo)
    B, K = 14, 2
    prev = torch.zeros((B, 512, 32, 1), device=dev)

    def post(obs, mask, pred_mono):                      # memory_nets.AcousticMem.forward on (pred_mono, previous memory)
        def sl(x):
            x = x.permute(0, 3, 1, 2)
            return x.reshape(x.size(0), x.size(1), 16, -1, x.size(3)).reshape(x.size(0), -1, 32, x.size(3))
        out = refiner(torch.cat((sl(pred_mono), sl(prev)), dim=1))
        return out.reshape(out.size(0), -1, 16, out.size(2), out.size(3)).reshape(out.size(0), 1, 512, out.size(3)).permute(0, 2, 3, 1)

    g = GraphedSeparation(pol.binSep_enc, pol.binSep_dec, dev, B, K, bin2mono_enc=pol.bin2mono_enc, bin2mono_dec=pol.bin2mono_dec, post=post)
    monos = torch.from_numpy(R.synth_monos(B, K, seed=4))
    rirs, lens = R.synth_rirs(B, K, L=16000, seed=14)
    tc = torch.from_numpy(np.random.default_rng(4).integers(0, 102, (B, 1)))
    g(monos, torch.from_numpy(rirs), torch.from_numpy(lens), tc)
    obs = g.obs.clone()
    with torch.no_grad():
        mask = pol.get_binSepMasks({"mixed_bin_audio_mag": obs, "target_class": tc.to(dev)})
        pred_mono = pol.convert_bin2mono(mask, mixed_audio=obs)
        mem = post(obs, mask, pred_mono)
    assert torch.equal(g.mask, mask) and torch.equal(g.pred_mono, pred_mono)
    assert g.pred_mono.shape == (B, 512, 32, 1) and g.post_out.shape == (B, 512, 32, 1)
    assert torch.allclose(g.post_out, mem, rtol=1e-4, atol=1e-5)


def test_graphed_rollout_step_with_native_acoustic_mem_and_stft_l2_against_oracle():
    """bin2mono prologue + second U-Net + AcousticMem + STFT_L2_distance replayed from ONE graph, checked
    against the oracle (render_dataset -> unet -> acoustic_mem_forward -> stft_l2_distance), not against eager CUDA."""
    from oracle import postproc as OP
    from move2hear_active_av_separation_b200.graphs import GraphedSeparation
    from move2hear_active_av_separation_b200.postproc import AcousticMem
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
    dev = torch.device("cuda", 0)
    sds = [*U.init_state(False, seed=3), *U.init_state(True, seed=4)]
    mods = [PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval(), PassiveSepEncCNN(True).to(dev).eval(), PassiveSepDecCNN(True).to(dev).eval()]
    for m, sd in zip(mods, sds):
        m.load_state_dict(sd)
    torch.manual_seed(7)
    mem = AcousticMem(use_ddppo=True).to(dev).eval()
    mem_sd = {k: v.detach().cpu() for k, v in mem.state_dict().items()}
    B, K = 14, 2
    g = GraphedSeparation(mods[0], mods[1], dev, B, K, bin2mono_enc=mods[2], bin2mono_dec=mods[3], acoustic_mem=mem, with_stft_l2=True)
    gen = torch.Generator().manual_seed(9)
    for seed in (0, 1):
        monos = R.synth_monos(B, K, seed=seed)
        rirs, lens = R.synth_rirs(B, K, L=16000, seed=20 + seed, ragged=True)
        tc = np.random.default_rng(seed).integers(0, 11, (B, 1))
        prev = torch.rand(B, 512, 32, 1, generator=gen)
        gt_bin = torch.stack([torch.rand(B, 512, 32, generator=gen), torch.rand(B, 512, 32, generator=gen) * 6 - 3,
                              torch.rand(B, 512, 32, generator=gen), torch.rand(B, 512, 32, generator=gen) * 6 - 3], dim=-1)
        gt_mono = torch.stack([torch.rand(B, 512, 32, generator=gen), torch.rand(B, 512, 32, generator=gen) * 6 - 3], dim=-1)
        g.prev_mem.copy_(prev)
        g.gt_bin_comps.copy_(gt_bin)
        g.gt_mono_comps.copy_(gt_mono)
        g(torch.from_numpy(monos), torch.from_numpy(rirs), torch.from_numpy(lens), torch.from_numpy(tc))
        torch.cuda.synchronize()
        obs = torch.from_numpy(np.stack([R.render_dataset(list(monos[b]), [R.rir_wav_layout(rirs[b, k], int(lens[b, k])) for k in range(K)], 0)[0]
                                         for b in range(B)]))
        with torch.no_grad():
            mask = U.separator_forward(sds[0], sds[1], obs, torch.from_numpy(tc))
            mono = U.separator_forward(sds[2], sds[3], mask, convert_bin2mono=True, mixed_audio=obs)
            want_mem = OP.acoustic_mem_forward(mem_sd, mono, prev, use_ddppo=True)
            wb, wm = OP.stft_l2_distance(obs, mask, gt_bin, mono, gt_mono)
            _, wmm = OP.stft_l2_distance(obs, mask, gt_bin, want_mem, gt_mono)
        rel = lambda a, b: float((a.cpu() - b).norm() / b.norm())
        assert rel(g.mask, mask) < 2e-2 and rel(g.pred_mono, mono) < 4e-2
        assert rel(g.pred_monoFromMem, want_mem) < 4e-2            # fp32 refiner on a bf16 U-Net's output
        # the refiner kernel itself, on the graph's own pred_mono: fp32-exact
        with torch.no_grad():
            assert rel(g.pred_monoFromMem, OP.acoustic_mem_forward(mem_sd, g.pred_mono.cpu(), prev, use_ddppo=True)) < 1e-5
            gb, gm = OP.stft_l2_distance(g.obs.cpu(), g.mask.cpu(), gt_bin, g.pred_mono.cpu(), gt_mono)
        assert torch.allclose(g.stft_l2["bin"].cpu(), gb, rtol=1e-5) and torch.allclose(g.stft_l2["mono"].cpu(), gm, rtol=1e-5)
        assert torch.allclose(g.stft_l2["bin"].cpu(), wb, rtol=5e-2) and torch.allclose(g.stft_l2["mono"].cpu(), wm, rtol=5e-2)
        assert torch.allclose(g.stft_l2["monoFromMem"].cpu(), wmm, rtol=5e-2)
