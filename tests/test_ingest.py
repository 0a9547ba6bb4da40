"""wav ingest + device-side clip / RIR caches against scipy.io.wavfile and the direct render."""
import io

import numpy as np
import pytest
from scipy.io import wavfile

from move2hear_active_av_separation_b200 import _lib


def wav_bytes(rate, arr):
    f = io.BytesIO()
    wavfile.write(f, rate, arr)
    return f.getvalue()


@pytest.mark.parametrize("dtype,shape", [(np.int16, (16000,)), (np.int16, (777, 2)), (np.int32, (100, 2)), (np.float32, (4001, 2)), (np.float32, (0, 2))])
def test_parse_wav_matches_scipy(dtype, shape):
    from move2hear_active_av_separation_b200.ingest import WAV_FLOAT32, WAV_PCM16, WAV_PCM32, parse_wav
    rng = np.random.default_rng(1)
    arr = (rng.standard_normal(shape) * 1000).astype(dtype)
    data = wav_bytes(16000, arr)
    info = parse_wav(data)
    sr, want = wavfile.read(io.BytesIO(data))
    assert info.sample_rate == sr == 16000
    assert info.frames == want.shape[0] and info.channels == (1 if want.ndim == 1 else want.shape[1])
    assert info.format == {np.dtype(np.int16): WAV_PCM16, np.dtype(np.int32): WAV_PCM32, np.dtype(np.float32): WAV_FLOAT32}[np.dtype(dtype)]
    n = want.size * want.dtype.itemsize
    got = np.frombuffer(data, np.uint8)[info.data_offset: info.data_offset + n].view(dtype).reshape(want.shape)
    np.testing.assert_array_equal(got, want)


def test_parse_wav_rejects_what_the_reference_treats_as_unreadable():
    from move2hear_active_av_separation_b200.ingest import parse_wav
    for bad in (b"", b"not a wav file at all", wav_bytes(16000, np.zeros((10, 2), np.int16))[:30],
                wav_bytes(16000, np.zeros((10, 2), np.uint8))):          # 8-bit PCM: wavfile.read would give uint8, never used by the reference
        with pytest.raises(ValueError):
            parse_wav(bad)


def test_row_cache_lru_and_pinning():
    from move2hear_active_av_separation_b200.ingest import RowCache
    c = RowCache(3)
    pinned = set()
    rows = [c.lookup(k, pinned)[0] for k in "abc"]
    assert sorted(rows) == [0, 1, 2] and c.misses == 3
    pinned = set()
    assert c.lookup("a", pinned) == (rows[0], True)              # "a" becomes most recent and is pinned for this batch
    row_d, hit = c.lookup("d", pinned)
    assert not hit and row_d == rows[1]                           # "b" was least recently used
    row_e, hit = c.lookup("e", pinned)
    assert not hit and row_e == rows[2]
    with pytest.raises(_lib.M2HError):
        c.lookup("f", pinned)                                     # every row belongs to the current batch


@pytest.mark.gpu
def test_cached_render_is_bit_identical_to_direct_render():
    import torch
    from move2hear_active_av_separation_b200.ingest import CachedAudioRenderer
    from move2hear_active_av_separation_b200.render import AudioRenderer
    rng = np.random.default_rng(7)
    N = 16000
    clips = {f"s{i}": np.round(3000 * rng.standard_normal(N)).astype(np.int16) for i in range(5)}
    clips["short"] = np.round(3000 * rng.standard_normal(9000)).astype(np.int16)          # zero-padded to N
    rirs = {}
    for i, (L, dt) in enumerate([(16000, np.float32), (4001, np.float32), (12000, np.int16), (7, np.int32), (15999, np.float32), (8000, np.float32)]):
        a = rng.standard_normal((L, 2)) * np.exp(-np.arange(L) / (L / 6 + 1))[:, None]
        rirs[("scene", 0, i, i + 1)] = (a * (0.05 if dt == np.float32 else 40)).astype(dt)
    rirs[("scene", 0, 99, 0)] = None                               # unreadable file -> zero RIR
    rirs[("scene", 0, 98, 0)] = np.zeros((0, 2), np.float32)       # empty file -> zero RIR
    as_bytes = {k: (wav_bytes(16000, v) if v is not None and i % 2 == 0 else v) for i, (k, v) in enumerate(rirs.items())}
    clip_bytes = {k: (wav_bytes(16000, v) if i % 2 else v) for i, (k, v) in enumerate(clips.items())}
    loads = {"clip": 0, "rir": 0}

    def clip_loader(name):
        loads["clip"] += 1
        return clip_bytes[name]

    def rir_loader(key):
        loads["rir"] += 1
        return as_bytes[key]

    rend = AudioRenderer()
    cached = CachedAudioRenderer(renderer=rend, rir_capacity=5, clip_capacity=4)         # smaller than the key sets (8 / 6): evictions happen
    ckeys, rkeys = list(clips), list(rirs)
    for step in range(10):
        B, K = 2, 2
        ck = [[ckeys[int(rng.integers(len(ckeys)))] for _ in range(K)] for _ in range(B)]
        rk = [[rkeys[int(rng.integers(len(rkeys)))] for _ in range(K)] for _ in range(B)]
        out = cached.render_keys(ck, rk, clip_loader, rir_loader, n_gt=1)
        for b in range(B):
            monos = [np.pad(clips[c], (0, N - len(clips[c]))) for c in ck[b]]
            rr = [rirs[r] if rirs[r] is not None else np.zeros((0, 2), np.float32) for r in rk[b]]
            (m, r, ln), _ = rend.pack_sources(monos, rr)
            want = rend.render(m, r, ln, n_gt=1)
            assert torch.equal(out["waves"][b], want["waves"][0])
            assert torch.equal(out["mix_mag"][b], want["mix_mag"][0])
            assert torch.equal(out["gt"][b], want["gt"][0])
    st = cached.stats()
    assert st["rir_hits"] + st["rir_misses"] == 40 and st["rir_misses"] == loads["rir"] and st["clip_misses"] == loads["clip"]
    assert st["rir_hits"] > 0 and st["rir_misses"] > len(rkeys) - 1      # hits and evictions both occurred


@pytest.mark.gpu
def test_cached_simulator_shim_equals_file_reading_shim(tmp_path):
    """get_current_mixed_bin_audio_mag_spec (wav files read every step, simulator_train.py:386-486) vs the cached variant."""
    import os
    from types import SimpleNamespace
    from move2hear_active_av_separation_b200 import render as MR
    rng = np.random.default_rng(3)
    mono_dir, rir_dir = tmp_path / "mono", tmp_path / "rirs"
    (rir_dir / "sceneA" / "90").mkdir(parents=True)
    mono_dir.mkdir()
    for name in ("dog", "speech"):
        wavfile.write(str(mono_dir / (name + ".wav")), 16000, np.round(3000 * rng.standard_normal(16000)).astype(np.int16))
    for recv in (4, 5):
        for src, L in ((11, 16000), (12, 9000)):
            a = (rng.standard_normal((L, 2)) * np.exp(-np.arange(L) / (L / 6))[:, None] * 0.05).astype(np.float32)
            wavfile.write(str(rir_dir / "sceneA" / "90" / f"{recv}_{src}.wav"), 16000, a)
    (rir_dir / "sceneA" / "90" / "6_11.wav").write_bytes(b"garbage")                    # unreadable -> zero RIR in both shims
    wavfile.write(str(rir_dir / "sceneA" / "90" / "6_12.wav"), 16000, np.zeros((0, 2), np.float32))     # empty -> zero RIR

    def make_sim():
        cfg = SimpleNamespace(AUDIO=SimpleNamespace(MONO_DIR=str(mono_dir), RIR_DIR=str(rir_dir), RIR_SAMPLING_RATE=16000,
                                                    GT_MONO_MAG_NORM=1.2, NORM_TYPE="l2"))
        return SimpleNamespace(config=cfg, _cached_mono_waveforms={}, _cached_mono_specs={}, _current_sound_names=["dog", "speech"],
                               _source_position_indices=[11, 12], current_scene_name="sceneA", azimuth_angle=90,
                               _receiver_position_index=4)

    a, b = make_sim(), make_sim()
    for recv in (4, 5, 4, 6, 5):
        a._receiver_position_index = b._receiver_position_index = recv
        obs_a = MR.get_current_mixed_bin_audio_mag_spec(a)
        obs_b = MR.get_current_mixed_bin_audio_mag_spec_cached(b)
        assert obs_a.dtype == obs_b.dtype == np.float16 and np.array_equal(obs_a.view(np.uint16), obs_b.view(np.uint16))
        for x, y in zip(a._gt_bin_audio_components, b._gt_bin_audio_components):
            assert np.array_equal(x.view(np.uint16), y.view(np.uint16))
        for x, y in zip(a._gt_mono_audio_components, b._gt_mono_audio_components):
            assert np.array_equal(x.view(np.uint16), y.view(np.uint16))
        assert np.array_equal(a._mixed_bin_audio_phase.view(np.uint16), b._mixed_bin_audio_phase.view(np.uint16))
    st = b._m2h_cached_renderer.stats()
    assert st["rir_misses"] == 6 and st["rir_hits"] == 4 and st["clip_misses"] == 2


def test_row_cache_invariants_random_traffic():
    """Model check of the LRU row cache: a key keeps its row until evicted, rows are never shared, a batch's rows are stable."""
    from move2hear_active_av_separation_b200.ingest import RowCache
    rng = np.random.default_rng(11)
    cap = 7
    c = RowCache(cap)
    model = {}                       # key -> row, as the cache should see it
    order = []                       # least recently used first
    for _ in range(400):
        batch = [int(k) for k in rng.integers(0, 15, size=int(rng.integers(1, cap + 1)))]
        pinned, rows_in_batch = set(), {}
        for k in batch:
            row, hit = c.lookup(k, pinned)
            assert hit == (k in model)
            if hit:
                assert row == model[k]
                order.remove(k)
            else:
                if len(model) == cap:                         # evict the least recently used key that this batch does not use
                    victim = next(v for v in order if model[v] not in pinned - {row})
                    assert model[victim] == row
                    del model[victim]
                    order.remove(victim)
                model[k] = row
            order.append(k)
            assert rows_in_batch.setdefault(k, row) == row       # stable inside the batch
        assert len(set(model.values())) == len(model) <= cap
    assert c.hits > 0 and c.misses > cap


def test_row_cache_forget_returns_the_row():
    """Transactional misses (CachedAudioRenderer.render_keys): a key whose row was never filled is unmapped again."""
    from move2hear_active_av_separation_b200.ingest import RowCache
    c = RowCache(2)
    pinned = set()
    row_a, _ = c.lookup("a", pinned)
    row_b, _ = c.lookup("b", pinned)
    c.forget("b")
    c.forget("never-inserted")
    assert "b" not in c.rows and c.rows == {"a": row_a} and c.free == [row_b]
    row_c, hit = c.lookup("c", set())
    assert not hit and row_c == row_b
    assert c.lookup("a", set()) == (row_a, True)
