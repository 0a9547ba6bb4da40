"""Consumers of the masks: STFT_L2_distance, AcousticMem, SI-SDR family.

CPU: ``oracle/postproc.py`` against ``tests/golden/postproc.npz`` = outputs of the reference's own
``eval_metrics.STFT_L2_distance`` / ``compute_waveform_quality`` and ``memory_nets.AcousticMem`` (unmodified, imported by
``oracle/gen_ref_golden.py::gen_postproc``).  GPU: the CUDA kernels (through the drop-in wrappers / C ABI) against both."""
import os

import numpy as np
import pytest
import torch

from oracle import postproc as OP
from oracle.gen_ref_golden import postproc_case_inputs, waveform_case_inputs

DB_TOL = 2e-3          # dB; the reference sums float32 samples, the kernel accumulates float64 moments


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "postproc.npz"))


def _sd(gold, tag):
    pre = f"mem_{tag}_sd."
    return {k[len(pre):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(pre)}


def test_oracle_stft_l2_matches_reference(gold):
    d = postproc_case_inputs()
    b, m = OP.stft_l2_distance(d["mixed_audio"], d["masks"], d["gt_bin_comps"], d["pred_mono"], d["gt_mono_comps"])
    np.testing.assert_array_equal(b.numpy(), gold["stft_l2_bin"])
    np.testing.assert_array_equal(m.numpy(), gold["stft_l2_mono"])


@pytest.mark.parametrize("tag", ["bn", "ddppo"])
def test_oracle_acoustic_mem_matches_reference(gold, tag):
    d = postproc_case_inputs()
    with torch.no_grad():
        y = OP.acoustic_mem_forward(_sd(gold, tag), d["pred_mono"], d["prev_mem"], use_ddppo=(tag == "ddppo"))
    assert tuple(y.shape) == (3, 512, 32, 1)
    np.testing.assert_allclose(y.numpy(), gold[f"mem_{tag}_out"], rtol=1e-5, atol=1e-6)


def test_oracle_waveform_quality_matches_reference(gold):
    names = [str(n) for n in gold["wq_names"]]
    q = OP.compute_waveform_quality(waveform_case_inputs(), names)
    np.testing.assert_allclose([q["mono"][k] for k in names], gold["wq_mono"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([q["monoFromMem"][k] for k in names], gold["wq_mem"], rtol=0, atol=1e-9)


def test_acoustic_mem_state_dict_layout_matches_reference(gold):
    from move2hear_active_av_separation_b200.postproc import AcousticMem
    for tag, ddppo in (("bn", False), ("ddppo", True)):
        mod = AcousticMem(use_ddppo=ddppo)
        want = _sd(gold, tag)
        assert sorted(mod.state_dict()) == sorted(want)
        assert all(tuple(mod.state_dict()[k].shape) == tuple(v.shape) for k, v in want.items())
        mod.load_state_dict(want)


# ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_stft_l2_kernel(gold):
    from move2hear_active_av_separation_b200.postproc import STFT_L2_distance
    d = {k: v.cuda() for k, v in postproc_case_inputs().items()}
    # the masks arrive as the decoder's permuted view, gt comps with all K sources concatenated (4K / 2K channels)
    masks = d["masks"].permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    gt_bin = torch.cat([d["gt_bin_comps"], torch.zeros_like(d["gt_bin_comps"])], dim=-1)
    gt_mono = torch.cat([d["gt_mono_comps"], torch.ones_like(d["gt_mono_comps"])], dim=-1)
    b, m = STFT_L2_distance(d["mixed_audio"], masks, gt_bin, d["pred_mono"], gt_mono)
    assert not b.is_cuda and tuple(b.shape) == (3, 1) and b.dtype == torch.float32
    np.testing.assert_allclose(b.numpy(), gold["stft_l2_bin"], rtol=2e-6)
    np.testing.assert_allclose(m.numpy(), gold["stft_l2_mono"], rtol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["bn", "ddppo"])
def test_acoustic_mem_kernel(gold, tag):
    from move2hear_active_av_separation_b200 import _lib
    from move2hear_active_av_separation_b200.postproc import AcousticMem
    d = postproc_case_inputs()
    mod = AcousticMem(use_ddppo=(tag == "ddppo")).cuda().eval()
    mod.load_state_dict(_sd(gold, tag))
    # pred_mono arrives as the bin2mono decoder's permuted view [B,512,32,1] with strides (16384, 32, 1, 16384)
    pm = d["pred_mono"].permute(0, 3, 1, 2).contiguous().cuda().permute(0, 2, 3, 1)
    with torch.no_grad():
        y = mod(pm, d["prev_mem"].cuda())
    assert tuple(y.shape) == (3, 512, 32, 1) and y.dtype == torch.float32
    np.testing.assert_allclose(y.cpu().numpy(), gold[f"mem_{tag}_out"], rtol=1e-4, atol=2e-5)
    # a larger, seeded batch against the oracle (rollout rank: B = 14)
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(14, 512, 32, 1, generator=g) * 2, torch.rand(14, 512, 32, 1, generator=g) * 2
    with torch.no_grad():
        got = mod(a.cuda(), b.cuda()).cpu()
        want = OP.acoustic_mem_forward(_sd(gold, tag), a, b, use_ddppo=(tag == "ddppo"))
    assert float((got - want).norm() / want.norm()) < 1e-5
    if tag == "bn":
        with pytest.raises(_lib.M2HError):
            mod.train()(a.cuda(), b.cuda())


@pytest.mark.gpu
def test_waveform_quality_on_gpu(gold):
    from move2hear_active_av_separation_b200.postproc import compute_waveform_quality, waveform_metrics
    names = [str(n) for n in gold["wq_names"]]
    q = compute_waveform_quality(waveform_case_inputs(), names)
    np.testing.assert_allclose([q["mono"][k] for k in names], gold["wq_mono"], rtol=0, atol=DB_TOL)
    np.testing.assert_allclose([q["monoFromMem"][k] for k in names], gold["wq_mem"], rtol=0, atol=DB_TOL)
    # the metric kernel alone on seeded waveforms, against the reference's numpy formulas (float64 inputs)
    rng = np.random.default_rng(2)
    B, L = 5, 16000
    ref = rng.standard_normal((B, L)).astype(np.float32) * 0.1 + 0.01
    est = (0.7 * ref + 0.05 * rng.standard_normal((B, L))).astype(np.float32)
    mix = (ref[:, None, :] * np.array([0.9, 0.6], np.float32)[None, :, None] + 0.2 * rng.standard_normal((B, 2, L))).astype(np.float32)
    got = waveform_metrics(torch.from_numpy(ref).cuda(), torch.from_numpy(est).cuda(), torch.from_numpy(mix).cuda())
    for i in range(B):
        want = OP.evaluate([ref[i:i + 1].astype(np.float64)], [est[i:i + 1].astype(np.float64)], [mix[i].astype(np.float64)])
        for k in ("si_sdr", "sd_sdr", "snr", "srr", "si_sdri", "sd_sdri", "snri", "si_sar"):
            assert abs(got[i][k] - want[k]) < 1e-6, (i, k, got[i][k], want[k])


@pytest.mark.gpu
def test_postproc_other_shapes():
    """B = 1, odd frame counts, permuted views: STFT_L2_distance against the oracle; AcousticMem at B = 1."""
    from move2hear_active_av_separation_b200.postproc import AcousticMem, STFT_L2_distance
    g = torch.Generator().manual_seed(17)
    for B, T in ((1, 32), (2, 7), (5, 13)):
        mixed = torch.log1p(torch.rand(B, 512, T, 2, generator=g) * 3)
        masks = torch.randn(B, 2, 512, T, generator=g).permute(0, 2, 3, 1)                   # the decoder's permuted view
        gt_bin = torch.stack([torch.rand(B, 512, T, generator=g), torch.rand(B, 512, T, generator=g) * 6 - 3,
                              torch.rand(B, 512, T, generator=g), torch.rand(B, 512, T, generator=g) * 6 - 3], dim=-1)
        mono = torch.rand(B, 1, 512, T, generator=g).permute(0, 2, 3, 1)
        gt_mono = torch.stack([torch.rand(B, 512, T, generator=g), torch.rand(B, 512, T, generator=g) * 6 - 3], dim=-1)
        wb, wm = OP.stft_l2_distance(mixed, masks, gt_bin, mono, gt_mono)
        b, m = STFT_L2_distance(mixed.cuda(), masks.cuda(), gt_bin.cuda(), mono.cuda(), gt_mono.cuda())
        assert tuple(b.shape) == (B, 1) and torch.allclose(b, wb, rtol=5e-6) and torch.allclose(m, wm, rtol=5e-6)
    torch.manual_seed(3)
    mem = AcousticMem(use_ddppo=False).cuda().eval()
    sd = {k: v.detach().cpu() for k, v in mem.state_dict().items()}
    a, p = torch.rand(1, 512, 32, 1, generator=g), torch.rand(1, 512, 32, 1, generator=g)
    with torch.no_grad():
        got, want = mem(a.cuda(), p.cuda()).cpu(), OP.acoustic_mem_forward(sd, a, p)
    assert float((got - want).norm() / want.norm()) < 1e-5
