"""Batched observation transport: render on the trainer's GPU from ``(sound ids, RIR ids)`` messages.

The reference renders every audio observation inside the env worker and ships it to the trainer as fp16 numpy arrays:
``Env.step`` -> sensors -> ``sim.get_current_mixed_bin_audio_mag_spec()`` ... -> pickle over a ``multiprocessing.Pipe``
(``audio_separation/common/env_utils.py:157-234, 348-381``) -> ``batch_obs`` / ``to_tensor`` widen to fp32 and upload
(``common/utils.py:66-97``) -> ``RolloutStorage.insert`` copies into ``observations[sensor][step + 1]``
(``common/rollout_storage.py:128-131``): ~0.75 MB per env and step (mag 64 KB + phase 64 KB + GT bin 256 KB + GT mono
128 KB as fp16, then 2x that as fp32 over PCIe), plus 15-36 ms of CPU render per env.

In the opt-in *deferred* mode the worker sends a ~100-byte :class:`AudioRenderRequest` instead; the trainer rank renders all
of its environments in ONE batched launch chain from device-resident clip / RIR pools (``ingest.CachedAudioRenderer``) and
writes the widened fp32 observations straight into the rollout storage's tensors:

* env side (bind in place of the simulator methods, like ``render.get_current_mixed_bin_audio_mag_spec``):
  :func:`get_current_mixed_bin_audio_mag_spec_deferred` and the three deferred getters - same call order and
  assertions as the reference (``simulator_train.py:363-384``), they return :class:`DeferredObservation` markers;
* trainer side: :class:`BatchedObservationRenderer` - :meth:`batch_obs` is a drop-in for ``common/utils.py::batch_obs``
  that resolves the markers, :meth:`insert_into` renders directly into ``rollouts.observations[...][step]``.

Values are identical to the per-process path: the same kernels run with the simulator dtypes (fp16 magnitude before log1p,
fp16 phases / GT) and ``to_tensor``'s fp16 -> fp32 widening is exact.
"""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass
from typing import Callable, Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ingest import CachedAudioRenderer

AUDIO_SENSORS = ("mixed_bin_audio_mag", "mixed_bin_audio_phase", "gt_bin_comps", "gt_mono_comps")


@dataclass(frozen=True)
class AudioRenderRequest:
    """Everything the render needs, as the simulator knows it at the step (simulator_train.py:395-406)."""
    sound_names: Tuple[str, ...]
    scene: str
    azimuth: int
    receiver: int
    sources: Tuple[int, ...]

    def rir_keys(self):
        return [(self.scene, self.azimuth, self.receiver, s) for s in self.sources]


@dataclass(frozen=True)
class DeferredObservation:
    """Placeholder an env worker puts into its observation dict; resolved by :class:`BatchedObservationRenderer`."""
    request: AudioRenderRequest
    field: str


# ---- env side ------------------------------------------------------------------------------------------------------------

def get_current_mixed_bin_audio_mag_spec_deferred(self):
    r"""Opt-in replacement for ``HabitatSimAudioEnabledTrain/Eval.get_current_mixed_bin_audio_mag_spec``: no file is read
    and nothing is rendered in the worker; the request is remembered for the three getters (which the sensors call after
    this method, as in the reference)."""
    req = AudioRenderRequest(tuple(self._current_sound_names), str(self.current_scene_name), int(self.azimuth_angle),
                             int(self._receiver_position_index), tuple(int(s) for s in self._source_position_indices))
    self._m2h_request = req
    # the reference's getters assert these lists are non-empty after the mag call (simulator_train.py:368, 377)
    self._gt_bin_audio_components = [DeferredObservation(req, "gt_bin_comps")]
    self._gt_mono_audio_components = [DeferredObservation(req, "gt_mono_comps")]
    self._mixed_bin_audio_phase = DeferredObservation(req, "mixed_bin_audio_phase")
    return DeferredObservation(req, "mixed_bin_audio_mag")


def get_current_gt_bin_audio_components_deferred(self):
    assert len(self._gt_bin_audio_components) != 0, " call this method after calling get_current_mixed_audio_mag_spec"
    return self._gt_bin_audio_components[0]


def get_current_gt_mono_audio_components_deferred(self):
    assert len(self._gt_mono_audio_components) != 0, "call this method after calling get_current_mixed_audio_mag_spec"
    return self._gt_mono_audio_components[0]


def get_current_mixed_bin_audio_phase_spec_deferred(self):
    return self._mixed_bin_audio_phase


# ---- trainer side --------------------------------------------------------------------------------------------------------

def to_tensor(v, device=None):
    """common/utils.py:66-72."""
    if torch.is_tensor(v):
        return v.to(device=device, dtype=torch.float)
    if isinstance(v, np.ndarray):
        return torch.from_numpy(v).to(device=device, dtype=torch.float)
    return torch.tensor(v, dtype=torch.float, device=device)


class BatchedObservationRenderer:
    """Resolves the deferred audio observations of one trainer rank's environments with one batched render per step."""

    def __init__(self, clip_loader: Callable[[str], object], rir_loader: Callable[[Hashable], object], device=None,
                 gt_mono_mag_norm: float = 0.0, cached: Optional[CachedAudioRenderer] = None, **cache_kwargs):
        self.cached = cached or CachedAudioRenderer(device=device, **cache_kwargs)
        self.device = self.cached.device
        self.clip_loader, self.rir_loader = clip_loader, rir_loader
        self.gt_mono_mag_norm = float(gt_mono_mag_norm)
        self._mono_specs: Dict[str, torch.Tensor] = {}          # sound name -> fp32 [512, T, 2] (fp16-rounded mag, phase): `_cached_mono_specs`
        self._clips_host: Dict[str, np.ndarray] = {}

    def _clip(self, name):
        if name not in self._clips_host:
            src = self.clip_loader(name)
            if not isinstance(src, np.ndarray):
                import io
                from scipy.io import wavfile
                _, src = wavfile.read(io.BytesIO(bytes(src)))
            self._clips_host[name] = src
        return self._clips_host[name]

    def _mono_spec(self, name):
        spec = self._mono_specs.get(name)
        if spec is None:
            mag, ph = self.cached.rend.mono_gt_spec(self._clip(name), self.gt_mono_mag_norm, want_phase=True)
            both = np.stack([mag, ph], axis=-1).astype(np.float16)                       # simulator_train.py:458-467
            spec = self._mono_specs[name] = torch.from_numpy(both).to(self.device).float()
        return spec

    def render(self, requests: Sequence[AudioRenderRequest]) -> Dict[str, torch.Tensor]:
        """``B`` requests (same number of sources) -> fp32 device tensors keyed by sensor uuid, batch-first."""
        B = len(requests)
        K = len(requests[0].sound_names)
        if any(len(r.sound_names) != K or len(r.sources) != K for r in requests):
            raise _lib.M2HError("all environments of a rank must have the same number of sources")
        out = self.cached.render_keys([list(r.sound_names) for r in requests], [r.rir_keys() for r in requests],
                                      self._clip, self.rir_loader, n_gt=K, sim_dtypes=True, want_phase=True)
        gt = out["gt"]                                                    # fp16 [B, K, 512, T, 4]
        T = gt.shape[3]
        gt_bin = gt.permute(0, 2, 3, 1, 4).reshape(B, 512, T, 4 * K).float()   # np.concatenate(axis=2) of the K [512,T,4] blocks
        gt_mono = torch.stack([torch.cat([self._mono_spec(n) for n in r.sound_names], dim=2) for r in requests])
        return {"mixed_bin_audio_mag": out["mix_mag"].float(), "mixed_bin_audio_phase": out["mix_phase"].float(),
                "gt_bin_comps": gt_bin, "gt_mono_comps": gt_mono}

    def _requests_of(self, observations: List[Dict]):
        reqs = []
        for obs in observations:
            req = next((v.request for v in obs.values() if isinstance(v, DeferredObservation)), None)
            if req is None:
                return None
            reqs.append(req)
        return reqs

    def batch_obs(self, observations: List[Dict], device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
        """Drop-in for ``common/utils.py::batch_obs`` (:75-97): deferred audio sensors are rendered here, every other
        sensor goes through ``to_tensor`` + ``stack`` like the reference."""
        batch = defaultdict(list)
        reqs = self._requests_of(observations)
        rendered = self.render(reqs) if reqs else {}
        for obs in observations:
            for sensor, v in obs.items():
                if not isinstance(v, DeferredObservation):
                    batch[sensor].append(to_tensor(v, device=device))
        result = {}
        for sensor in observations[0]:
            v0 = observations[0][sensor]
            if isinstance(v0, DeferredObservation):
                t = rendered[v0.field]
                result[sensor] = t if device is None else t.to(device)
            else:
                result[sensor] = torch.stack(batch[sensor], dim=0)
        return result

    def insert_into(self, observations: List[Dict], storage_observations: Dict[str, torch.Tensor], step: int) -> None:
        """``RolloutStorage.insert``'s observation loop (rollout_storage.py:128-131): ``storage[sensor][step].copy_(...)``
        with the rendered tensors coming straight from the render's device buffers."""
        batch = self.batch_obs(observations, device=None)
        for sensor, t in batch.items():
            if sensor in storage_observations:
                storage_observations[sensor][step].copy_(t, non_blocking=True)
