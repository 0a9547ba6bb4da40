"""CPU: the render kernels' bodies, run phase-by-phase on the host (tests/emu), against the oracle.

This checks every index computation and the fp32 numerics of csrc/render_conv.cuh / render_stft.cuh without a
GPU.  The emulation is test infrastructure: the product library never links it.
"""
import os

import numpy as np
import pytest

from oracle import render as R
from oracle.gen_golden import RENDER_CASES, render_case_inputs
from util import f16_close, load_emu, max_rel, ptr, rel_l2


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def run_conv(emu, monos, rp, ln, n_ctas=2):
    K, N = monos.shape
    mono = np.ascontiguousarray(monos[None])
    rir = np.ascontiguousarray(rp[None])
    rl = np.ascontiguousarray(ln[None].astype(np.int32))
    waves = np.full((1, K, 2, N), 777, np.int16)
    emu.emu_rir_convolve(ptr(mono), ptr(rir), ptr(rl), 1, K, N, rp.shape[-1], ptr(waves), n_ctas)
    return waves


@pytest.mark.parametrize("case", RENDER_CASES, ids=[c[0] for c in RENDER_CASES])
def test_emulated_kernels_match_golden(case, emu, golden_dir):
    name, K, lens, seed = case
    g = np.load(os.path.join(golden_dir, f"render_{name}.npz"))
    monos, rirs, rp, ln = render_case_inputs(K, lens, seed)
    N = monos.shape[1]
    T = 1 + (N - 1) // 512
    waves = run_conv(emu, monos, rp, ln)
    want = g["source_waves_i16"]
    d = waves[0].astype(np.int32) - want.astype(np.int32)
    assert np.abs(d).max() <= 1
    assert np.count_nonzero(d) <= 2e-3 * d.size            # 1-LSB flips on near-ties only
    assert rel_l2(waves[0], want) < 1e-4 and max_rel(waves[0], want) < 1e-4

    # dataset variant (fp32): mixture log1p magnitude + target GT magnitude
    mix = np.zeros((1, 512, T, 2), np.float32)
    gt = np.zeros((1, 1, 512, T, 2), np.float32)
    emu.emu_stft_waves(ptr(waves), 1, K, N, 1, 0, 0, 0, ptr(mix), None, ptr(gt), 1)
    assert rel_l2(mix[0], g["ds_mixed_logmag_f32"]) < 1e-4 and max_rel(mix[0], g["ds_mixed_logmag_f32"]) < 1e-4
    if np.abs(g["ds_gt_bin_mag_f32"]).max() > 0:
        assert rel_l2(gt[0, 0], g["ds_gt_bin_mag_f32"]) < 1e-4 and max_rel(gt[0, 0], g["ds_gt_bin_mag_f32"]) < 1e-4
    else:
        assert np.all(gt == 0)

    # simulator variant (fp16 + phases + all-source GT)
    mixh = np.zeros((1, 512, T, 2), np.float16)
    ph = np.zeros((1, 512, T, 2), np.float16)
    gth = np.zeros((1, K, 512, T, 4), np.float16)
    emu.emu_stft_waves(ptr(waves), 1, K, N, K, 1, 1, 0, ptr(mixh), ptr(ph), ptr(gth), 3)
    ok, nbad, maxulp = f16_close(mixh[0], g["sim_mixed_mag_f16"])
    assert ok, (nbad, maxulp)
    gm = g["sim_gt_bin_f16"].astype(np.float32)
    for ear in (0, 2):
        za = gth[0][..., ear].astype(np.float32) * np.exp(1j * gth[0][..., ear + 1].astype(np.float32))
        zb = gm[..., ear] * np.exp(1j * gm[..., ear + 1])
        assert np.linalg.norm(za - zb) <= 1e-3 * max(np.linalg.norm(zb), 1e-30)      # fp16 storage: 2^-11 per component
    zp = np.exp(1j * ph[0].astype(np.float32))
    zo = np.exp(1j * g["sim_mixed_phase_f16"].astype(np.float32))
    w = np.expm1(g["sim_mixed_mag_f16"].astype(np.float32))
    assert np.linalg.norm((zp - zo) * w) <= 2e-3 * np.linalg.norm(w)


def test_emulated_f16_observation_is_bit_exact_given_oracle_waves(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "render_k2_L16000.npz"))
    waves = np.ascontiguousarray(g["source_waves_i16"][None])
    mixh = np.zeros((1, 512, 32, 2), np.float16)
    emu.emu_stft_waves(ptr(waves), 1, 2, 16000, 0, 0, 1, 0, ptr(mixh), None, None, 1)
    neq = np.count_nonzero(mixh[0].view(np.uint16) != g["sim_mixed_mag_f16"].view(np.uint16))
    ok, nbad, maxulp = f16_close(mixh[0], g["sim_mixed_mag_f16"])
    assert neq <= 33 and ok, (neq, nbad, maxulp)         # <= 0.1 % differ: rounding ties of the fp16 magnitude cast


def test_emulated_stft_f32_mono(emu):
    mono = R.synth_monos(1, 1, seed=4)[0, 0]
    wave = np.ascontiguousarray((mono.astype("float32") / 32768)[None, None])
    mag = np.zeros((1, 512, 32, 1), np.float32)
    ph = np.zeros((1, 512, 32, 1), np.float32)
    emu.emu_stft_f32(ptr(wave), 1, 1, 16000, ptr(mag), ptr(ph))
    D = R.stft_librosa(wave[0, 0])
    assert rel_l2(mag[0, :, :, 0], np.abs(D)) < 1e-5
    assert rel_l2(mag[0, :, :, 0] * np.exp(1j * ph[0, :, :, 0]), D) < 1e-5


def test_emulated_conv_wraps_like_numpy(emu):
    # out-of-range values wrap modulo 2^16 (numpy astype("int16")), SURVEY hard part 3
    N = 16000
    monos = np.zeros((1, N), np.int16)
    monos[0, :4] = [20000, -20000, 16384, 100]
    rp = np.zeros((1, 2, 1), np.float32)
    rp[0, 0, 0], rp[0, 1, 0] = 2.0, -2.0
    waves = run_conv(emu, monos, rp, np.array([1]), n_ctas=1)
    np.testing.assert_array_equal(waves[0, 0, 0, :4], [-25536, 25536, -32768, 200])
    np.testing.assert_array_equal(waves[0, 0, 1, :4], [25536, -25536, -32768, -200])
    assert not waves[0, 0, :, 4:].any()


def test_emulated_conv_short_signal(emu):
    # N < 16000 and odd RIR lengths exercise the window-offset arithmetic
    rng = np.random.default_rng(5)
    # (5000, 12002) / (16000, 48002): even row length with an odd first tap -> the packed-load variant's one-tap shift
    for N, L in ((1024, 333), (5000, 12001), (16384, 32767), (5000, 12002), (16000, 48002), (16384, 32768)):
        mono = np.round(3000 * rng.standard_normal((1, N))).astype(np.int16)
        h = (rng.standard_normal((1, 2, L)) * 0.02).astype(np.float32)
        waves = run_conv(emu, mono, h, np.array([L]), n_ctas=1)
        want = R.convolve_quantise(mono[0], np.ascontiguousarray(h[0].T)) * 32768
        d = waves[0, 0].astype(np.int32) - want.astype(np.int32)
        # near-tie flips grow with the transform's content: 0.05 % (L = 333) .. 0.33 % (32767 taps, N = 16384); the oracle's own
        # complex64 RIR spectrum contributes flips of the same order against exact float64 convolution
        assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= 4e-3 * d.size, (N, L)


def test_unet_slab_schedule_host_emulation(tmp_path):
    """Item tables + weight-tile packing of the slab U-Net kernel, replayed on the CPU against direct (de)convolutions."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "emu_unet_slab")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-w", "-I", os.path.join(root, "include"),
                           "-o", exe, os.path.join(root, "tests", "emu", "emu_unet_slab.cu")])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("N", [16384, 5004, 512, 1536])
def test_emulated_stft_other_lengths(emu, N):
    # N a multiple of 512 (frame count 1 + (N-1)//512), N not a multiple of 8 (scalar staging path), shortest signals
    rng = np.random.default_rng(N)
    K = 3
    waves = np.round(2000 * rng.standard_normal((1, K, 2, N))).astype(np.int16)
    T = R.n_frames(N)
    mix = np.zeros((1, 512, T, 2), np.float32)
    gt = np.zeros((1, K, 512, T, 2), np.float32)
    emu.emu_stft_waves(ptr(waves), 1, K, N, K, 0, 0, 1, ptr(mix), None, ptr(gt), 2)
    mixed = (waves[0].astype(np.float32) * np.float32(1 / 32768)).sum(0) / np.float32(K)
    for ear in range(2):
        want = np.abs(R.stft_librosa(mixed[ear]))
        assert want.shape == (512, T)
        assert rel_l2(mix[0, :, :, ear], want) < 1e-4 and max_rel(mix[0, :, :, ear], want) < 1e-4
        for k in range(K):
            wk = np.abs(R.stft_librosa(waves[0, k, ear].astype(np.float32) * np.float32(1 / 32768)))
            assert rel_l2(gt[0, k, :, :, ear], wk) < 1e-4


@pytest.mark.parametrize("C,T,length", [(1, 32, 16000), (2, 32, 16000), (2, 7, 3000), (1, 33, 20000)])
def test_emulated_istft_matches_librosa_restatement(emu, C, T, length):
    """m2h_istft's kernel body (Bluestein inverse DFT of 1022 points, window, overlap-add) vs oracle.render.istft_librosa."""
    rng = np.random.default_rng(100 * C + T)
    B = 2
    mag = np.abs(rng.standard_normal((B, 512, T, C))).astype(np.float32)
    ph = rng.uniform(-np.pi, np.pi, (B, 512, T, C)).astype(np.float32)
    out = np.full((B, C, length), 7.0, np.float32)
    emu.emu_istft(ptr(mag), ptr(ph), B, C, T, length, ptr(out), 3)
    for b in range(B):
        for c in range(C):
            want = R.istft_librosa(mag[b, :, :, c] * np.exp(1j * ph[b, :, :, c]), length=length)
            # (1, 33, 20000) runs past the last frame's centre: its final samples are divided by a vanishing window sum
            # (values ~1e3, ill-conditioned in any precision) and dominate both norms
            assert rel_l2(out[b, c], want) < 1e-4 and max_rel(out[b, c], want) < 1e-4
