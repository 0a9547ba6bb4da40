"""GPU parity: the CUDA render path (through the C ABI) against the oracle and the committed golden vectors."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import render as R
from oracle.gen_golden import RENDER_CASES, render_case_inputs
from util import f16_close, max_rel, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rend():
    from move2hear_active_av_separation_b200.render import AudioRenderer
    return AudioRenderer(torch.device("cuda", 0))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("case", RENDER_CASES, ids=[c[0] for c in RENDER_CASES])
def test_render_matches_golden(case, rend, golden_dir):
    name, K, lens, seed = case
    g = np.load(os.path.join(golden_dir, f"render_{name}.npz"))
    monos, rirs, rp, ln = render_case_inputs(K, lens, seed)
    mono_d, rir_d, len_d = dev(monos[None]), dev(rp[None]), dev(ln[None])
    # dataset variant: fp32, target GT magnitude only
    out = rend.render(mono_d, rir_d, len_d, n_gt=1)
    waves = out["waves"][0].cpu().numpy()
    want = g["source_waves_i16"]
    d = waves.astype(np.int32) - want.astype(np.int32)
    assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= 2e-3 * d.size
    # tolerance: 1e-4 relative, normwise; element-wise the floor is ONE quantisation step (checked above):
    # on low-amplitude cases (k2_short peaks at ~340 LSB) a single legitimate tie flip is already 3e-3 of the maximum
    assert rel_l2(waves, want) < 1e-4 and max_rel(waves, want) <= max(1e-4, 1.0 / np.abs(want).max())
    mix = out["mix_mag"][0].cpu().numpy()
    assert mix.dtype == np.float32 and mix.shape == (512, 32, 2)
    want_mix = g["ds_mixed_logmag_f32"]
    if np.count_nonzero(d):
        # a legitimate 1-LSB tie flip (bounded above) moves low-amplitude spectra by more than 1e-4 of their maximum: the
        # spectrogram is then held to 1e-4 against the oracle's own STFT chain applied to the SAME integer waves
        mixed = np.zeros((2, waves.shape[-1]), np.float32)
        for k in range(K):
            mixed += waves[k].astype(np.float32) * np.float32(1 / 32768)
        mixed /= np.float32(K)
        want_mix = np.log1p(np.stack([np.abs(R.stft_librosa(mixed[ch])) for ch in range(2)], axis=-1).astype(np.float32))
        assert rel_l2(want_mix, g["ds_mixed_logmag_f32"]) < 1e-3
    assert rel_l2(mix, want_mix) < 1e-4 and max_rel(mix, want_mix) < 1e-4
    gt = out["gt"][0, 0].cpu().numpy()
    if np.abs(g["ds_gt_bin_mag_f32"]).max() > 0:
        assert rel_l2(gt, g["ds_gt_bin_mag_f32"]) < 1e-4 and max_rel(gt, g["ds_gt_bin_mag_f32"]) < 1e-4
    else:
        assert not gt.any()
    # simulator variant: fp16, phases, all-source GT
    out = rend.render(mono_d, rir_d, len_d, n_gt=K, sim_dtypes=True, want_phase=True)
    mixh = out["mix_mag"][0].cpu().numpy()
    assert mixh.dtype == np.float16
    ok, nbad, maxulp = f16_close(mixh, g["sim_mixed_mag_f16"])
    assert ok, (nbad, maxulp)
    gth = out["gt"][0].cpu().numpy().astype(np.float32)
    gm = g["sim_gt_bin_f16"].astype(np.float32)
    for ear in (0, 2):
        za = gth[..., ear] * np.exp(1j * gth[..., ear + 1])
        zb = gm[..., ear] * np.exp(1j * gm[..., ear + 1])
        assert np.linalg.norm(za - zb) <= 1e-3 * max(np.linalg.norm(zb), 1e-30)
    ph = out["mix_phase"][0].cpu().numpy().astype(np.float32)
    w = np.expm1(g["sim_mixed_mag_f16"].astype(np.float32))
    assert np.linalg.norm((np.exp(1j * ph) - np.exp(1j * g["sim_mixed_phase_f16"].astype(np.float32))) * w) <= 2e-3 * np.linalg.norm(w)


def test_f16_observation_bit_exact_given_oracle_waves(rend, golden_dir):
    g = np.load(os.path.join(golden_dir, "render_k2_L16000.npz"))
    mix, _, _ = rend.stft_waves(dev(g["source_waves_i16"][None]), sim_dtypes=True)
    mix = mix[0].cpu().numpy()
    neq = np.count_nonzero(mix.view(np.uint16) != g["sim_mixed_mag_f16"].view(np.uint16))
    ok, nbad, maxulp = f16_close(mix, g["sim_mixed_mag_f16"])
    assert neq <= 33 and ok, (neq, nbad, maxulp)


def test_batch_against_oracle_seeded(rend):
    # ragged batch; the oracle finishes B=6 in about a second
    B, K = 6, 2
    monos = R.synth_monos(B, K, seed=0)
    rirs, lens = R.synth_rirs(B, K, L=16000, seed=1, ragged=True)
    out = rend.render(dev(monos), dev(rirs), dev(lens), n_gt=1)
    mix = out["mix_mag"].cpu().numpy()
    gt = out["gt"].cpu().numpy()
    for b in range(B):
        rl = [R.rir_wav_layout(rirs[b, k], int(lens[b, k])) for k in range(K)]
        ds = R.render_dataset(list(monos[b]), rl, 0)
        assert rel_l2(mix[b], ds[0]) < 1e-4 and max_rel(mix[b], ds[0]) < 1e-4
        assert rel_l2(gt[b, 0], ds[1][0]) < 1e-4


def test_full_size_properties(rend):
    """BASELINE config 2 size (B = 1024): size-independent properties instead of a full oracle run."""
    B, K, N = 1024, 2, 16000
    monos = R.synth_monos(8, K, seed=0)
    rirs, lens = R.synth_rirs(8, K, seed=1)
    rep = B // 8
    mono_d = dev(monos).repeat(rep, 1, 1)
    rir_d = dev(rirs).repeat(rep, 1, 1, 1)
    len_d = dev(lens).repeat(rep, 1)
    out = rend.render(mono_d, rir_d, len_d)
    waves, mix = out["waves"], out["mix_mag"]
    # (1) replicas are bit-identical wherever they sit in the batch (no cross-item interference, persistent loop ok)
    assert torch.equal(waves.view(rep, 8, K, 2, N), waves[:8].expand(rep, 8, K, 2, N).contiguous().view(rep, 8, K, 2, N))
    assert torch.equal(mix.view(rep, 8, 512, 32, 2)[-1], mix[:8])
    # (2) spot-check three items against the oracle
    for b in (0, 5, 7):
        rl = [R.rir_wav_layout(rirs[b, k], int(lens[b, k])) for k in range(K)]
        ds = R.render_dataset(list(monos[b]), rl, 0)
        assert rel_l2(mix[B - 8 + b].cpu().numpy(), ds[0]) < 1e-4
    # (3) swapping the ears of the RIR swaps the ears: exactly for the quantised waves (ears are convolved
    #     independently), to fp32 rounding for the spectrogram (the ears share one complex transform)
    out2 = rend.render(mono_d[:16], rir_d[:16].flip(2), len_d[:16])
    assert torch.equal(out2["waves"], waves[:16].flip(2))
    assert rel_l2(out2["mix_mag"].cpu().numpy(), mix[:16].flip(3).cpu().numpy()) < 1e-5
    # (4) permuting the sources leaves the mixture unchanged (integer sum before the divide)
    out3 = rend.render(mono_d[:16].flip(1), rir_d[:16].flip(1), len_d[:16].flip(1))
    assert torch.equal(out3["mix_mag"], mix[:16])
    # (5) a unit impulse RIR (L = 1, gain 1) returns the clip itself
    imp = torch.zeros(4, K, 2, 1, device="cuda")
    imp[..., 0] = 1.0
    w = rend.rir_convolve(mono_d[:4], imp, torch.ones(4, K, dtype=torch.int32, device="cuda"))
    assert torch.equal(w, mono_d[:4, :, None, :].expand(4, K, 2, N))


def test_edge_cases(rend):
    # empty batch
    e = rend.render(torch.empty(0, 2, 16000, dtype=torch.int16, device="cuda"),
                    torch.empty(0, 2, 2, 16, device="cuda"), torch.empty(0, 2, dtype=torch.int32, device="cuda"))
    assert e["mix_mag"].shape == (0, 512, 32, 2)
    # all RIRs unreadable -> zeros -> log1p(0) = 0
    z = rend.render(dev(R.synth_monos(2, 2, seed=3)), torch.zeros(2, 2, 2, 8, device="cuda"),
                    torch.zeros(2, 2, dtype=torch.int32, device="cuda"))
    assert not z["waves"].any() and not z["mix_mag"].any()
    # maximum supported clip length
    from move2hear_active_av_separation_b200 import M2HError
    with pytest.raises(M2HError):
        rend.render(torch.zeros(1, 1, 16385, dtype=torch.int16, device="cuda"), torch.zeros(1, 1, 2, 8, device="cuda"),
                    torch.ones(1, 1, dtype=torch.int32, device="cuda"))
    with pytest.raises(M2HError):
        rend.rir_convolve(torch.zeros(1, 1, 16000, dtype=torch.int16), torch.zeros(1, 1, 2, 8), torch.ones(1, 1, dtype=torch.int32))


def test_stft_f32_against_oracle(rend):
    mono = R.synth_monos(1, 2, seed=4)[0]
    wave = torch.from_numpy(mono.astype("float32") / 32768).cuda()[None]
    mag, ph = rend.stft_f32(wave, want_phase=True)
    for c in range(2):
        D = R.stft_librosa(wave[0, c].cpu().numpy())
        assert rel_l2(mag[0, :, :, c].cpu().numpy(), np.abs(D)) < 1e-5
        assert rel_l2(mag[0, :, :, c].cpu().numpy() * np.exp(1j * ph[0, :, :, c].cpu().numpy()), D) < 1e-5


def test_sim_shim_side_effects(rend, golden_dir):
    """Drop-in for simulator_train.py:386-486: return value + the three side-effect fields."""
    from move2hear_active_av_separation_b200 import render as MR
    g = np.load(os.path.join(golden_dir, "render_k2_ragged.npz"))
    monos, rirs, _, _ = render_case_inputs(2, [7999, 4000], 12)
    sim = SimpleNamespace(_cached_mono_specs={}, _gt_bin_audio_components=[], _gt_mono_audio_components=[])
    obs = MR.render_sim_arrays(sim, list(monos), rirs, ["a", "b"], 1.2, rend)
    assert obs.dtype == np.float16 and obs.shape == (512, 32, 2)
    assert f16_close(obs, g["sim_mixed_mag_f16"])[0]
    assert len(sim._gt_bin_audio_components) == 2 and sim._gt_bin_audio_components[0].shape == (512, 32, 4)
    assert np.concatenate(sim._gt_bin_audio_components, axis=2).shape == (512, 32, 8)      # task.py GtBin sensor
    gm = np.stack(sim._gt_mono_audio_components)
    assert gm.dtype == np.float16 and gm.shape == (2, 512, 32, 2)
    assert rel_l2(gm[..., 0].astype(np.float32), g["sim_gt_mono_f16"][..., 0].astype(np.float32)) < 1e-3
    assert sim._mixed_bin_audio_phase.shape == (512, 32, 2) and set(sim._cached_mono_specs) == {"a", "b"}


def test_dataset_shim(rend, golden_dir):
    """Drop-in for dataset.py:162-228."""
    from move2hear_active_av_separation_b200 import render as MR
    g = np.load(os.path.join(golden_dir, "render_k2_L16000.npz"))
    monos, rirs, _, _ = render_case_inputs(2, [16000, 16000], 11)
    ds = SimpleNamespace(_gt_mono_mag_cache={})
    mixed, gt_bin, gt_mono, tc = MR.compute_audiospects_arrays(ds, list(monos), rirs, ["x.wav", "y.wav"], 3, 1.2, rend)
    assert mixed.dtype == np.float32 and rel_l2(mixed, g["ds_mixed_logmag_f32"]) < 1e-4
    assert rel_l2(gt_bin[0], g["ds_gt_bin_mag_f32"]) < 1e-4
    assert gt_mono[0].shape == (512, 32, 1) and rel_l2(gt_mono[0], g["ds_gt_mono_mag_f32"]) < 1e-4
    assert tc.dtype == np.int64 and tc.tolist() == [3] and "x.wav" in ds._gt_mono_mag_cache


def test_istft_matches_librosa_restatement(rend):
    """m2h_istft (eval_metrics.istft) against the oracle's librosa.istft restatement, incl. the drop-in."""
    import torch
    from move2hear_active_av_separation_b200 import render as MR
    rng = np.random.default_rng(9)
    y = (rng.standard_normal((2, 16000)) * 0.1).astype(np.float32)
    D = np.stack([R.stft_librosa(y[c]) for c in range(2)], axis=-1)                 # [512, 32, 2] complex64
    mag, ph = np.abs(D).astype(np.float32), np.angle(D).astype(np.float32)
    out = rend.istft(torch.from_numpy(mag[None]).cuda(), torch.from_numpy(ph[None]).cuda())[0].cpu().numpy()
    for c in range(2):
        want = R.istft_librosa(mag[..., c] * np.exp(1j * ph[..., c]))
        assert rel_l2(out[c], want) < 1e-4 and max_rel(out[c], want) < 1e-4
    sig = MR.istft(mag[..., 0], ph[..., 0], mag[..., 1], ph[..., 1], rend=rend)
    want = R.istft_eval(mag[..., 0], ph[..., 0], mag[..., 1], ph[..., 1])
    assert len(sig) == 2 and sig[0].dtype == np.float32 and sig[0].shape == (16000,)
    for a, b in zip(sig, want):
        assert rel_l2(a, b) < 1e-4
    mono = MR.istft(mag[..., 0], ph[..., 0], rend=rend)
    assert len(mono) == 1 and rel_l2(mono[0], want[0]) < 1e-4
    # batch of random (inconsistent) spectra, odd frame count, longer output
    B, T = 5, 13
    m = np.abs(rng.standard_normal((B, 512, T, 2))).astype(np.float32)
    p = rng.uniform(-np.pi, np.pi, (B, 512, T, 2)).astype(np.float32)
    o = rend.istft(torch.from_numpy(m).cuda(), torch.from_numpy(p).cuda(), length=7000).cpu().numpy()
    for b in range(B):
        for c in range(2):
            want = R.istft_librosa(m[b, :, :, c] * np.exp(1j * p[b, :, :, c]), length=7000)
            assert rel_l2(o[b, c], want) < 1e-4


def test_conv_and_stft_random_shapes(rend):
    """Seeded sweep over clip lengths, RIR lengths (shorter than, equal to and up to 3 x the clip), source counts, ragged
    lengths inside a batch, even / odd row lengths (packed and scalar kernel variants, odd first taps)."""
    rng = np.random.default_rng(2024)
    shapes = [(16000, 16000, 2, 3), (16000, 15999, 2, 2), (16000, 48002, 3, 2), (16000, 32001, 1, 2), (8192, 8192, 2, 2),
              (5000, 12002, 2, 3), (5001, 777, 2, 2), (16384, 32768, 1, 2), (12000, 1, 2, 2), (16000, 40000, 4, 1),
              # both sides of the dispatch between the F = 24576 kernel (Lmax <= 16384) and the F = 32768 kernel, largest clip
              (16384, 16384, 2, 2), (16384, 16386, 2, 2), (16000, 16383, 1, 3), (1536, 16384, 2, 2), (513, 2, 3, 2)]
    for N, Lmax, K, B in shapes:
        mono = np.round(2500 * rng.standard_normal((B, K, N))).astype(np.int16)
        lens = rng.integers(max(1, Lmax // 3), Lmax + 1, size=(B, K)).astype(np.int32)
        lens[0, 0] = Lmax
        decay = np.exp(-np.arange(Lmax) / (Lmax / 6 + 1)).astype(np.float32)
        rir = (rng.standard_normal((B, K, 2, Lmax)).astype(np.float32) * decay * 0.03)
        for b in range(B):
            for k in range(K):
                rir[b, k, :, lens[b, k]:] = 7.0            # garbage beyond the valid taps must never be read
        out = rend.render(dev(mono), dev(rir), dev(lens), n_gt=K, raw_mag=True)
        waves = out["waves"].cpu().numpy()
        for b in range(B):
            for k in range(K):
                h = np.ascontiguousarray(rir[b, k, :, :lens[b, k]].T)
                want = (R.convolve_quantise(mono[b, k], h) * 32768).astype(np.int32)
                d = waves[b, k].astype(np.int32) - want
                assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= 4e-3 * d.size, (N, Lmax, K, b, k)
            if N >= 512:
                mixed = (waves[b].astype(np.float32) * np.float32(1 / 32768)).sum(0) / np.float32(K)
                for ear in range(2):
                    want = np.abs(R.stft_librosa(mixed[ear]))
                    got = out["mix_mag"][b, :, :, ear].cpu().numpy()
                    assert got.shape == want.shape and rel_l2(got, want) < 1e-4, (N, Lmax, K, b, ear)


def test_stft_gt_only_launch_equals_combined_launch(rend):
    """M2H_RENDER_GT_ONLY: the per-source spectrograms from their own launch (a training step runs it on a side stream) are
    bit-identical to those of the combined mixture + GT launch."""
    mono = R.synth_monos(5, 3, seed=77)
    rir, ln = R.synth_rirs(5, 3, L=9000, seed=78, ragged=True)
    waves = rend.rir_convolve(torch.from_numpy(mono).cuda(), torch.from_numpy(rir).cuda(), torch.from_numpy(ln).cuda())
    for sim in (False, True):
        mix, _, gt = rend.stft_waves(waves, 2, sim_dtypes=sim, want_phase=sim)
        none, _, gt2 = rend.stft_waves(waves, 2, sim_dtypes=sim, want_phase=sim, gt_only=True)
        assert none is None and torch.equal(gt, gt2)
        mix0, _, _ = rend.stft_waves(waves, 0, sim_dtypes=sim)
        assert torch.equal(mix, mix0)
    assert rend.stft_waves(waves, 0, gt_only=True) == (None, None, None)
