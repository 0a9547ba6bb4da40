"""GPU parity: the tcgen05 U-Net (through the C ABI / drop-in modules) against the oracle and the reference golden."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unet as U
from oracle.gen_golden import unet_case_inputs
from util import rel_l2

pytestmark = pytest.mark.gpu

# bf16 operands, fp32 accumulation: mask rel-L2 <= 2e-2 vs the fp32 torch oracle
TOL = 2e-2


def build(bin2mono, seed):
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
    enc_sd, dec_sd = U.init_state(bin2mono, seed=seed)
    enc, dec = PassiveSepEncCNN(bin2mono).cuda().eval(), PassiveSepDecCNN(bin2mono).cuda().eval()
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)
    return enc, dec, enc_sd, dec_sd


def test_state_dict_names_match_reference_layout():
    from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
    pol = Move2HearPassiveWoMemoryPolicy()
    sd = pol.state_dict()
    assert len(sd) == 124 and sum(v.numel() for v in sd.values()) == 33478148
    assert sd["binSep_enc.passive_sep_encoder.cnn.0.0.weight"].shape == (64, 33, 4, 4)
    assert sd["bin2mono_dec.passive_sep_decoder.cnn.5.0.bias"].shape == (16,)
    assert "binSep_dec.passive_sep_decoder.cnn.1.1.num_batches_tracked" in sd


@pytest.mark.parametrize("bin2mono", [False, True])
def test_matches_reference_golden(bin2mono, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_bin2mono.npz" if bin2mono else "unet_binsep.npz"))
    enc, dec, _, _ = build(bin2mono, 3 + int(bin2mono))
    obs, tc = unet_case_inputs(3, seed=5)
    with torch.no_grad():
        if bin2mono:
            bott, skips = enc(torch.from_numpy(g["in_masks"]).cuda(), mixed_audio=obs.cuda())
        else:
            bott, skips = enc({"mixed_bin_audio_mag": obs.cuda(), "target_class": tc.cuda()})
        mask = dec(bott, skips)
    assert tuple(bott.shape) == (3, 512)
    assert [tuple(s.shape) for s in skips] == [(3, 512, 2, 2), (3, 256, 4, 4), (3, 128, 8, 8), (3, 64, 16, 16)]
    assert rel_l2(skips[3].float().cpu().numpy(), g["eval_skip3"].astype(np.float32)) < TOL        # first layer
    assert rel_l2(bott.cpu().numpy(), g["eval_bottleneck"]) < TOL                                  # whole encoder
    assert tuple(mask.shape) == g["eval_mask"].shape and list(mask.stride()) == list(g["eval_mask_strides"])
    assert mask.dtype == torch.float32
    assert rel_l2(mask.cpu().numpy(), g["eval_mask"]) < TOL


@pytest.mark.parametrize("B", [1, 14, 130])
def test_layerwise_against_oracle(B):
    enc, dec, enc_sd, dec_sd = build(False, 7)
    obs, tc = unet_case_inputs(B, seed=B)
    with torch.no_grad():
        bott, skips = enc({"mixed_bin_audio_mag": obs.cuda(), "target_class": tc.cuda()})
        mask = dec(bott, skips)
        wb, ws = U.encoder_forward(enc_sd, obs, tc)
        wm = U.decoder_forward(dec_sd, wb, ws)
    for got, want in zip(skips[::-1], ws[::-1]):                     # shallow -> deep localises a broken layer
        assert rel_l2(got.float().cpu().numpy(), want.numpy()) < TOL, tuple(want.shape)
    assert rel_l2(bott.cpu().numpy(), wb.numpy()) < TOL
    # decoder intermediates live in the lower channel half of the concat buffers
    w = bott._m2h_ws
    x = wb.reshape(B, 512, 1, 1)
    with torch.no_grad():
        for i, (idx, hw, c) in enumerate(((4, 2, 512), (3, 4, 256), (2, 8, 128), (1, 16, 64))):
            if i > 0:
                x = torch.cat((x, ws[i - 1]), dim=1)
            x = F.relu(F.batch_norm(F.conv_transpose2d(x, dec_sd[f"cnn.{i}.0.weight"], None, stride=2, padding=1),
                                    dec_sd[f"cnn.{i}.1.running_mean"], dec_sd[f"cnn.{i}.1.running_var"],
                                    dec_sd[f"cnn.{i}.1.weight"], dec_sd[f"cnn.{i}.1.bias"], False, 0.1, 1e-5))
            got = w.view(idx, (B, hw, hw, 2 * c))[..., :c].permute(0, 3, 1, 2).float().cpu().numpy()
            assert rel_l2(got, x.numpy()) < TOL, ("decoder layer", i)
    assert rel_l2(mask.cpu().numpy(), wm.numpy()) < TOL


def test_policy_composition_and_loss():
    """get_binSepMasks -> convert_bin2mono -> L1 losses (passive_trainer.py:228-236, 269-276) within 1e-2 relative."""
    from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
    torch.manual_seed(0)
    pol = Move2HearPassiveWoMemoryPolicy().cuda().eval()
    sd = {k: v.clone() for k, v in pol.state_dict().items()}
    for k in sd:                                                       # non-trivial BN statistics
        if k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand_like(sd[k])
        if k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn_like(sd[k])
    pol.load_state_dict(sd)
    B = 6
    obs, tc = unet_case_inputs(B, seed=21)
    g = torch.Generator().manual_seed(1)
    gt_bin, gt_mono = torch.rand(B, 512, 32, 2, generator=g), torch.rand(B, 512, 32, 1, generator=g)
    with torch.no_grad():
        masks = pol.get_binSepMasks({"mixed_bin_audio_mag": obs.cuda(), "target_class": tc.cuda()})
        mono = pol.convert_bin2mono(masks.detach(), mixed_audio=obs.cuda())
        bl, ml = U.supervised_losses(obs.cuda(), masks, gt_bin.cuda(), mono, gt_mono.cuda())
    cpu = {k: v.cpu() for k, v in sd.items()}
    sub = lambda pre: {k[len(pre):]: v for k, v in cpu.items() if k.startswith(pre)}
    with torch.no_grad():
        wm = U.separator_forward(sub("binSep_enc.passive_sep_encoder."), sub("binSep_dec.passive_sep_decoder."), obs, tc)
        wmono = U.separator_forward(sub("bin2mono_enc.passive_sep_encoder."), sub("bin2mono_dec.passive_sep_decoder."),
                                    wm, None, True, obs)
        wbl, wml = U.supervised_losses(obs, wm, gt_bin, wmono, gt_mono)
    assert tuple(mono.shape) == (B, 512, 32, 1)
    assert rel_l2(masks.cpu().numpy(), wm.numpy()) < TOL and rel_l2(mono.cpu().numpy(), wmono.numpy()) < 2 * TOL
    assert abs(float(bl) - float(wbl)) / float(wbl) < 1e-2 and abs(float(ml) - float(wml)) / float(wml) < 1e-2


def test_decoder_accepts_foreign_features_and_repacks_after_weight_update():
    enc, dec, enc_sd, dec_sd = build(False, 9)
    obs, tc = unet_case_inputs(2, seed=2)
    with torch.no_grad():
        wb, ws = U.encoder_forward(enc_sd, obs, tc)
        mask = dec(wb.cuda(), [s.cuda() for s in ws])               # plain fp32 NCHW tensors, not workspace views
        want = U.decoder_forward(dec_sd, wb, ws)
        assert rel_l2(mask.cpu().numpy(), want.numpy()) < TOL
        dec.cnn[5][0].bias.add_(1.0)                                  # in-place update must invalidate the packed weights
        mask2 = dec(wb.cuda(), [s.cuda() for s in ws])
    assert abs(float((mask2 - mask).mean()) - 1.0) < 1e-3


def test_no_fallbacks():
    from move2hear_active_av_separation_b200 import M2HError
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepEncCNN
    obs, tc = unet_case_inputs(1, seed=2)
    with pytest.raises(M2HError):           # CPU module / CPU tensors: no silent torch path in either mode
        PassiveSepEncCNN().eval()({"mixed_bin_audio_mag": obs, "target_class": tc})
    with pytest.raises(M2HError):
        PassiveSepEncCNN().train()({"mixed_bin_audio_mag": obs, "target_class": tc})
    with pytest.raises(M2HError):           # train mode updates fp32 running statistics in place: half-precision modules are refused
        PassiveSepEncCNN().cuda().half().train()({"mixed_bin_audio_mag": obs.cuda(), "target_class": tc.cuda()})


def test_workspace_pool_keeps_value_semantics():
    """Workspaces are reused (their zero borders are written once), but never while a previous result is alive."""
    enc, dec, _, _ = build(False, 11)
    obs_a, tc_a = unet_case_inputs(5, seed=21)
    obs_b, tc_b = unet_case_inputs(5, seed=22)
    with torch.no_grad():
        bott_a, skips_a = enc({"mixed_bin_audio_mag": obs_a.cuda(), "target_class": tc_a.cuda()})
        keep = [s.clone() for s in skips_a]
        ptr_a = bott_a._m2h_ws.buf.data_ptr()
        bott_b, skips_b = enc({"mixed_bin_audio_mag": obs_b.cuda(), "target_class": tc_b.cuda()})    # a's outputs are alive
        assert bott_b._m2h_ws.buf.data_ptr() != ptr_a
        for s, k in zip(skips_a, keep):
            assert torch.equal(s, k)
        mask_a = dec(bott_a, skips_a).clone()
        mask_b = dec(bott_b, skips_b).clone()
        del bott_a, skips_a, bott_b, skips_b, s
        import gc
        gc.collect()        # a lease caught in a reference cycle (seen under compute-sanitizer) is released by the collector
        # both workspaces are free again: the next passes reuse them and reproduce the results bit for bit
        for _ in range(3):
            bott, skips = enc({"mixed_bin_audio_mag": obs_a.cuda(), "target_class": tc_a.cuda()})
            assert len(enc._ws_pool.items) == 2
            assert torch.equal(dec(bott, skips), mask_a)
            del bott, skips
            bott, skips = enc({"mixed_bin_audio_mag": obs_b.cuda(), "target_class": tc_b.cuda()})
            assert torch.equal(dec(bott, skips), mask_b)
            del bott, skips
