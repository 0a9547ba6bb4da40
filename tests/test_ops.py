"""The ``m2h::`` torch.library operators (north_star: "C-ABI extension registered with PyTorch")."""
import numpy as np
import pytest
import torch

import move2hear_active_av_separation_b200.ops  # noqa: F401  (registers the namespace)
from oracle import render as R
from oracle import unet as U
from oracle.gen_golden import unet_case_inputs

OPS = ["rir_convolve", "stft_waves", "render_mixture", "istft", "unet_forward", "bin2mono_forward", "stft_l2_distance", "acoustic_mem", "stft_waves_x0", "unet_forward_x0"]


def test_namespace_is_registered_cuda_only():
    for name in OPS:
        assert hasattr(torch.ops.m2h, name), name
    # no CPU kernel is registered: the dispatcher refuses CPU tensors (there is no fallback to take)
    with pytest.raises(NotImplementedError):
        torch.ops.m2h.rir_convolve(torch.zeros(1, 2, 16000, dtype=torch.int16), torch.zeros(1, 2, 2, 16000), torch.zeros(1, 2, dtype=torch.int32))
    with pytest.raises(NotImplementedError):
        torch.ops.m2h.unet_forward(torch.zeros(1, 512, 32, 2), torch.zeros(1, 1), torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8))


def test_fake_kernels_give_shapes_and_dtypes():
    with torch.device("meta"):
        mono, rir, ln = torch.zeros(3, 2, 16000, dtype=torch.int16), torch.zeros(3, 2, 2, 16000), torch.zeros(3, 2, dtype=torch.int32)
        mix, phase, gt, waves = torch.ops.m2h.render_mixture(mono, rir, ln, 2, True, True)
        assert mix.shape == (3, 512, 32, 2) and mix.dtype == torch.float16 and phase.shape == mix.shape
        assert gt.shape == (3, 2, 512, 32, 4) and waves.shape == (3, 2, 2, 16000) and waves.dtype == torch.int16
        mask = torch.ops.m2h.unet_forward(torch.zeros(3, 512, 32, 2), torch.zeros(3, 1), torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8))
        assert mask.shape == (3, 2, 512, 32)
        assert torch.ops.m2h.istft(torch.zeros(2, 512, 32, 2), torch.zeros(2, 512, 32, 2), 16000).shape == (2, 2, 16000)
        x0 = torch.ops.m2h.stft_waves_x0(waves)
        assert x0.shape == (3, 18, 18, 128) and x0.dtype == torch.bfloat16
        assert torch.ops.m2h.unet_forward_x0(x0, torch.zeros(3, 1), torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8)).shape == (3, 2, 512, 32)


@pytest.mark.gpu
def test_operators_against_oracle():
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
    dev = torch.device("cuda", 0)
    B, K = 3, 2
    monos = R.synth_monos(B, K, seed=5)
    rirs, lens = R.synth_rirs(B, K, L=16000, seed=6, ragged=True)
    mix, phase, gt, waves = torch.ops.m2h.render_mixture(torch.from_numpy(monos).to(dev), torch.from_numpy(rirs).to(dev),
                                                         torch.from_numpy(lens).to(dev), 1, False, False)
    assert phase.numel() == 0 and gt.shape == (B, 1, 512, 32, 2)
    want = np.stack([R.render_dataset(list(monos[b]), [R.rir_wav_layout(rirs[b, k], int(lens[b, k])) for k in range(K)], 0)[0] for b in range(B)])
    assert np.linalg.norm(mix.cpu().numpy() - want) < 1e-4 * np.linalg.norm(want)
    # the separator operators on the packed weights of the drop-in modules
    enc_sd, dec_sd = U.init_state(False, seed=3)
    enc, dec = PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval()
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)
    obs, tc = unet_case_inputs(B, seed=8)
    mask = torch.ops.m2h.unet_forward(obs.to(dev), tc.to(dev), enc._pack(dev), dec._pack(dev)).permute(0, 2, 3, 1)
    want_mask = U.separator_forward(enc_sd, dec_sd, obs, tc)
    assert float((mask.cpu() - want_mask).norm() / want_mask.norm()) < 2e-2
    with torch.no_grad():
        assert torch.equal(mask, dec(*enc({"mixed_bin_audio_mag": obs.to(dev), "target_class": tc.to(dev)})))
    # fused hand-over: STFT -> first-layer operand -> separator, bit-identical to the route through the fp32 observation
    tc3 = tc.to(dev)
    m_two_step = torch.ops.m2h.unet_forward(mix, tc3, enc._pack(dev), dec._pack(dev)).clone()
    m_fused = torch.ops.m2h.unet_forward_x0(torch.ops.m2h.stft_waves_x0(waves), tc3, enc._pack(dev), dec._pack(dev))
    assert torch.equal(m_two_step, m_fused)
    enc2_sd, dec2_sd = U.init_state(True, seed=4)
    enc2, dec2 = PassiveSepEncCNN(True).to(dev).eval(), PassiveSepDecCNN(True).to(dev).eval()
    enc2.load_state_dict(enc2_sd)
    dec2.load_state_dict(dec2_sd)
    mono = torch.ops.m2h.bin2mono_forward(mask, obs.to(dev), enc2._pack(dev), dec2._pack(dev)).permute(0, 2, 3, 1)
    want_mono = U.separator_forward(enc2_sd, dec2_sd, want_mask, convert_bin2mono=True, mixed_audio=obs)
    assert float((mono.cpu() - want_mono).norm() / want_mono.norm()) < 4e-2
    # opaque launches on the current stream: capturable
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        torch.ops.m2h.unet_forward(obs.to(dev), tc.to(dev), enc._pack(dev), dec._pack(dev))
    torch.cuda.current_stream(dev).wait_stream(s)
    o, t = obs.to(dev), tc.to(dev)
    with torch.cuda.graph(g):
        m2 = torch.ops.m2h.unet_forward(o, t, enc._packed, dec._packed)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(m2.permute(0, 2, 3, 1), mask)
