"""Batched observation transport: deferred env-side observations resolved by one batched GPU render per step,
against what the reference's path delivers (oracle.render_sim -> fp16 numpy -> ``to_tensor`` -> stack -> storage copy)."""
import pickle
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import render as R
from util import f16_close


def _sim(names, recv, sources):
    return SimpleNamespace(_current_sound_names=list(names), current_scene_name="sceneA", azimuth_angle=90, _receiver_position_index=recv,
                           _source_position_indices=list(sources), _gt_bin_audio_components=[], _gt_mono_audio_components=[],
                           _mixed_bin_audio_phase=None)


def test_env_side_markers_follow_the_reference_call_protocol():
    from move2hear_active_av_separation_b200 import transport as TR
    sim = _sim(["dog", "speech"], 4, [11, 12])
    with pytest.raises(AssertionError):                      # simulator_train.py:368: getters only after the mag call
        TR.get_current_gt_bin_audio_components_deferred(sim)
    mag = TR.get_current_mixed_bin_audio_mag_spec_deferred(sim)
    obs = {"mixed_bin_audio_mag": mag, "mixed_bin_audio_phase": TR.get_current_mixed_bin_audio_phase_spec_deferred(sim),
           "gt_bin_comps": TR.get_current_gt_bin_audio_components_deferred(sim),
           "gt_mono_comps": TR.get_current_gt_mono_audio_components_deferred(sim)}
    assert [v.field for v in obs.values()] == list(obs)
    assert mag.request.rir_keys() == [("sceneA", 90, 4, 11), ("sceneA", 90, 4, 12)]
    wire = pickle.dumps(obs)                                  # what crosses the worker's Pipe (env_utils.py:186-188)
    assert len(wire) < 1024 and pickle.loads(wire) == obs     # vs ~0.5 MB of fp16 spectrograms per env and step


@pytest.mark.gpu
def test_batched_render_into_rollout_storage_matches_reference_path():
    from move2hear_active_av_separation_b200 import transport as TR
    dev = torch.device("cuda", 0)
    n_envs, K, T = 14, 2, 32
    names = [f"s{i}" for i in range(5)]
    clips = {n: R.synth_monos(1, 1, seed=300 + i)[0, 0] for i, n in enumerate(names)}
    rirs = {}
    for recv in range(4):
        for src in (11, 12, 13):
            L = 4000 + 3000 * ((recv + src) % 5)
            rp, _ = R.synth_rirs(1, 1, L=L, seed=400 + 10 * recv + src)
            rirs[("sceneA", 90, recv, src)] = R.rir_wav_layout(rp[0, 0], L)
    rirs[("sceneA", 90, 3, 13)] = np.zeros((0, 2), np.float32)              # empty file -> zero RIR
    bor = TR.BatchedObservationRenderer(clips.__getitem__, rirs.__getitem__, device=dev, gt_mono_mag_norm=1.2, rir_capacity=64, clip_capacity=8)
    storage = {"mixed_bin_audio_mag": torch.zeros(3, n_envs, 512, T, 2, device=dev), "mixed_bin_audio_phase": torch.zeros(3, n_envs, 512, T, 2, device=dev),
               "gt_bin_comps": torch.zeros(3, n_envs, 512, T, 4 * K, device=dev), "gt_mono_comps": torch.zeros(3, n_envs, 512, T, 2 * K, device=dev),
               "target_class": torch.zeros(3, n_envs, 1, device=dev)}
    rng = np.random.default_rng(0)
    for step in (1, 2):
        sims, observations = [], []
        for e in range(n_envs):
            pair = rng.choice(len(names), size=K, replace=False)
            sim = _sim([names[i] for i in pair], int(rng.integers(4)), [int(s) for s in rng.choice([11, 12, 13], size=K, replace=False)])
            observations.append({"mixed_bin_audio_mag": TR.get_current_mixed_bin_audio_mag_spec_deferred(sim),
                                 "mixed_bin_audio_phase": TR.get_current_mixed_bin_audio_phase_spec_deferred(sim),
                                 "gt_bin_comps": TR.get_current_gt_bin_audio_components_deferred(sim),
                                 "gt_mono_comps": TR.get_current_gt_mono_audio_components_deferred(sim),
                                 "target_class": np.array([e % 11])})
            sims.append(sim)
        observations = pickle.loads(pickle.dumps(observations))           # the Pipe
        bor.insert_into(observations, storage, step)
        torch.cuda.synchronize()
        for e, sim in enumerate(sims):
            want = R.render_sim([clips[n] for n in sim._current_sound_names],
                                [rirs[("sceneA", 90, sim._receiver_position_index, s)] for s in sim._source_position_indices], gt_mono_mag_norm=1.2)
            got_mag = storage["mixed_bin_audio_mag"][step, e].cpu().numpy()
            assert got_mag.dtype == np.float32
            ok, nbad, maxulp = f16_close(got_mag.astype(np.float16), want["mixed_bin_audio_mag"])
            assert ok and np.array_equal(got_mag, got_mag.astype(np.float16).astype(np.float32)), (nbad, maxulp)   # fp16 values widened exactly
            want_bin = np.concatenate(want["gt_bin_audio_components"], axis=2).astype(np.float32)          # simulator_train.py:369
            got_bin = storage["gt_bin_comps"][step, e].cpu().numpy()
            assert got_bin.shape == want_bin.shape == (512, T, 4 * K)
            for c in range(0, 4 * K, 2):
                za, zb = got_bin[..., c] * np.exp(1j * got_bin[..., c + 1]), want_bin[..., c] * np.exp(1j * want_bin[..., c + 1])
                assert np.linalg.norm(za - zb) <= 1e-3 * max(np.linalg.norm(zb), 1e-30)
            want_mono = np.concatenate(want["gt_mono_audio_components"], axis=2).astype(np.float32)        # :378
            got_mono = storage["gt_mono_comps"][step, e].cpu().numpy()
            assert got_mono.shape == want_mono.shape == (512, T, 2 * K)
            for c in range(0, 2 * K, 2):
                assert np.linalg.norm(got_mono[..., c] - want_mono[..., c]) <= 1e-3 * np.linalg.norm(want_mono[..., c])
            ph, wph = storage["mixed_bin_audio_phase"][step, e].cpu().numpy(), want["mixed_bin_audio_phase"].astype(np.float32)
            w = np.expm1(want["mixed_bin_audio_mag"].astype(np.float32))
            assert np.linalg.norm((np.exp(1j * ph) - np.exp(1j * wph)) * w) <= 2e-3 * np.linalg.norm(w)
            assert float(storage["target_class"][step, e, 0]) == e % 11
    st = bor.cached.stats()
    assert st["rir_hits"] > 0 and st["clip_hits"] > st["clip_misses"]
    assert not storage["mixed_bin_audio_mag"][0].any()                      # untouched step


@pytest.mark.gpu
def test_batched_render_three_sources_and_mixed_source_counts_rejected():
    from move2hear_active_av_separation_b200 import _lib
    from move2hear_active_av_separation_b200 import transport as TR
    dev = torch.device("cuda", 0)
    clips = {f"s{i}": R.synth_monos(1, 1, seed=500 + i)[0, 0] for i in range(3)}
    rirs = {}
    for src, L in ((1, 12000), (2, 3000), (3, 16000)):
        rp, _ = R.synth_rirs(1, 1, L=L, seed=600 + src)
        rirs[("sc", 0, 0, src)] = R.rir_wav_layout(rp[0, 0], L)
    bor = TR.BatchedObservationRenderer(clips.__getitem__, rirs.__getitem__, device=dev, gt_mono_mag_norm=0.0, rir_capacity=8, clip_capacity=4)
    req3 = TR.AudioRenderRequest(("s0", "s1", "s2"), "sc", 0, 0, (1, 2, 3))
    out = bor.render([req3, req3])
    want = R.render_sim([clips["s0"], clips["s1"], clips["s2"]], [rirs[("sc", 0, 0, s)] for s in (1, 2, 3)], gt_mono_mag_norm=0.0)
    assert tuple(out["gt_bin_comps"].shape) == (2, 512, 32, 12) and tuple(out["gt_mono_comps"].shape) == (2, 512, 32, 6)
    ok, nbad, maxulp = f16_close(out["mixed_bin_audio_mag"][1].cpu().numpy().astype(np.float16), want["mixed_bin_audio_mag"])
    assert ok, (nbad, maxulp)
    gm = np.concatenate(want["gt_mono_audio_components"], axis=2).astype(np.float32)
    got = out["gt_mono_comps"][0].cpu().numpy()
    for c in range(0, 6, 2):                                           # GT_MONO_MAG_NORM == 0: the simulator keeps raw magnitudes
        assert np.linalg.norm(got[..., c] - gm[..., c]) <= 1e-3 * np.linalg.norm(gm[..., c])
    req2 = TR.AudioRenderRequest(("s0", "s1"), "sc", 0, 0, (1, 2))
    with pytest.raises(_lib.M2HError):
        bor.render([req3, req2])
