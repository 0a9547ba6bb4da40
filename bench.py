#!/usr/bin/env python
"""Headline benchmark: separated mixtures / second through (RIR conv + quantise + mix + STFT + U-Net).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic mixtures: BASELINE.json configs[1]
(B = 1024 mixtures per GPU, K = 2 sources x 1 s @ 16 kHz, 16000-tap binaural RIRs, seeded as SURVEY.md 8d) rendered
to log1p-magnitude spectrograms, then the binaural passive-separator U-Net (random init, eval-mode BN) predicting
the ratio masks.  Work shards by mixture across GPUs (no data-path collective, weak scaling).  Prints ONE JSON line.

``--impl reference`` times the reference's own CPU arithmetic (scipy fftconvolve + librosa-faithful STFT + the
torch-fp32 U-Net restated in oracle/) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "separated mixtures/sec (RIR conv+STFT+U-Net)"
UNIT = "mixtures/s"
B_PER_GPU = 1024
K_SRC = 2
N_SAMPLES = 16000
L_TAPS = 16000
N_DISTINCT = 64               # distinct seeded mixtures, tiled to the batch (identical for every arm)

# algorithmic bytes per mixture: every input read once, every output written once
BYTES_CONV = 2 * K_SRC * N_SAMPLES + 8 * K_SRC * L_TAPS + 2 * K_SRC * 2 * N_SAMPLES        # mono + RIR in, int16 waves out
BYTES_STFT = 2 * K_SRC * 2 * N_SAMPLES + 4 * 512 * 32 * 2                                   # int16 waves in, fp32 log-mag out
BYTES_RENDER = 2 * K_SRC * N_SAMPLES + 8 * K_SRC * L_TAPS + 4 * 512 * 32 * 2                # 451,072 B
FLOPS_UNET = 220.7e6                                                                         # binSep fwd, nominal 2*MACs
BYTES_STFT_X0 = 2 * K_SRC * 2 * N_SAMPLES + 2 * 128 * 16 * 16                               # int16 waves in, bf16 first-layer operand out
# kernels of this library launched per timed step: m2h_rir_conv24_kernel, m2h_stft_kernel (writes the separator's bf16 first-layer
# operand: no pack_input_kernel, no bf16_to_f32_kernel for the bottleneck), 3 x unet_slab_kernel (enc0, dec3, dec4 + head),
# 7 x unet_gemm_tma_kernel (enc1-4, dec0-2) - the ncu launch list of the same command shows these 12 and nothing else
OUR_LAUNCHES_PER_STEP = 12


def workload_config(extra=None):
    cfg = {"workload": "configs[1] render (B=1024/GPU, K=2, N=16000, L=16000, seeds 0/1) + binSep U-Net fwd (eval BN)",
           "batch_per_gpu": B_PER_GPU, "sources": K_SRC, "samples": N_SAMPLES, "rir_taps": L_TAPS,
           "l2": "inputs+intermediates per step (0.9 GB) exceed the 126 MB L2; no artificial flush",
           "warmup_note": "W warm-up steps plus 40 more untimed steps of the same load so that nvidia-smi clock samples bracket the ~30 ms timed region"}
    if extra:
        cfg.update(extra)
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def synth_host_inputs():
    import numpy as np
    from oracle import render as R
    monos = R.synth_monos(N_DISTINCT, K_SRC, N_SAMPLES, seed=0)
    rirs, lens = R.synth_rirs(N_DISTINCT, K_SRC, L_TAPS, seed=1)
    tc = (np.arange(N_DISTINCT) % 11).astype(np.int64).reshape(-1, 1)
    return monos, rirs, lens, tc


# ------------------------------------------------------------------------------------------------
# CPU reference path (oracle) - the checker doubling as the timed "reference arm"
# ------------------------------------------------------------------------------------------------
def _pool_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"


def _cpu_render_one(args):
    from oracle import render as R
    monos, rirs, lens = args
    rl = [R.rir_wav_layout(rirs[k], int(lens[k])) for k in range(len(monos))]
    return R.render_dataset(list(monos), rl, 0)[0]


class CpuReference:
    """scipy/numpy render on a process pool (one mixture per task, like the reference's env / DataLoader workers)
    followed by the torch-fp32 U-Net on all host threads."""

    def __init__(self):
        import multiprocessing as mp
        import torch
        from oracle import unet as U
        self.cores = len(os.sched_getaffinity(0))
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_pool_init)
        torch.set_num_threads(self.cores)
        self.enc_sd, self.dec_sd = U.init_state(False, seed=3)
        self.U = U
        self.pool.map(_cpu_render_one, [])          # spin the workers up

    def step(self, monos, rirs, lens, tc):
        import numpy as np
        import torch
        obs = self.pool.map(_cpu_render_one, [(monos[i], rirs[i], lens[i]) for i in range(len(monos))])
        obs = torch.from_numpy(np.stack(obs))
        with torch.no_grad():
            mask = self.U.separator_forward(self.enc_sd, self.dec_sd, obs, torch.from_numpy(tc))
        return mask

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample_size(cores):
    """Mixtures per CPU step of the in-line `cpu_baseline` leg: a few tasks per worker, bounded so that a step stays around a second."""
    return max(16, min(256, 4 * cores))


def run_reference(args):
    """Reference arm: the reference's CPU arithmetic on the box's host cores, on OUR arm's configuration - every step is
    the full 1024-mixture batch of configs[1] (about 1.4 s per step on 16 cores, so the default 20 + 3 steps take ~35 s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    monos, rirs, lens, tc = synth_host_inputs()
    ref = CpuReference()
    sample = B_PER_GPU
    sl = [i % N_DISTINCT for i in range(sample)]
    for _ in range(max(args.warmup, 1)):
        ref.step(monos[sl], rirs[sl], lens[sl], tc[sl])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(monos[sl], rirs[sl], lens[sl], tc[sl])
    dt = time.perf_counter() - t0
    ref.close()
    value = sample * args.steps / dt
    world = args.gpus
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
            "config": workload_config({"global_batch": world * B_PER_GPU, "parallelism": f"shard-by-mixture x{world}, no data-path collective"}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": "port",
                             "sample": f"{sample} mixtures/step (one full per-GPU batch of the same seeded workload; the host does not grow with N), "
                                       f"scipy.fftconvolve + restated librosa STFT on a {ref.cores}-process pool, then torch-fp32 U-Net on {ref.cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def _collect(self, lo, hi):
        sm, mx, reasons = [], None, set()
        for t, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not (lo <= t <= hi):
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return sm, mx, reasons

    def stop(self, t0, t1, t_load0=None):
        """Samples taken inside the timed region [t0, t1]; the region is ~30 ms, so when fewer than two samples fall inside
        it the window is widened to the continuous load around it (warm-up steps from ``t_load0`` on, 50 ms after)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = self._collect(t0, t1)
        window = "timed region"
        if len(sm) < 2:
            sm, mx, reasons = self._collect(t0 if t_load0 is None else t_load0, t1 + 0.05)
            window = "timed region + the warm-up steps before it (same load)"
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------
# BASELINE configs[3] as written: 14 env PROCESSES per GPU, each calling the B = 1 drop-in shim once per env step
# (common/env_utils.py:236-269 forkserver workers; task.py:85-86 -> sim.get_current_mixed_bin_audio_mag_spec)
# ------------------------------------------------------------------------------------------------
def _rollout_worker(idx, local, n_steps, barrier, out_q):
    import numpy as np
    import torch
    from types import SimpleNamespace
    torch.cuda.set_device(local)
    from move2hear_active_av_separation_b200 import render as MR
    from oracle import render as R
    monos = R.synth_monos(1, K_SRC, N_SAMPLES, seed=1000 + idx)[0]
    rp, ln = R.synth_rirs(8, K_SRC, L_TAPS, seed=2000 + idx, ragged=True)
    rirs = [[R.rir_wav_layout(rp[j, k], int(ln[j, k])) for k in range(K_SRC)] for j in range(8)]
    sim = SimpleNamespace(_cached_mono_waveforms={}, _cached_mono_specs={}, config=SimpleNamespace(AUDIO=SimpleNamespace(NORM_TYPE="l2")))
    names = ["a", "b"]
    for j in range(3):                                   # context, library load, mono-spec cache
        MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)
    barrier.wait()
    t0 = time.perf_counter()
    for j in range(n_steps):
        obs = MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)     # returns host fp16 [512,32,2] + the 3 side-effect fields
    dt = time.perf_counter() - t0
    barrier.wait()
    out_q.put((idx, dt, float(np.asarray(obs, dtype=np.float32).sum())))


def run_rollout_processes(local, n_procs=14, n_steps=40):
    """env-steps/s of `n_procs` processes sharing GPU `local` (one CUDA context each), each rendering one mixture per step."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n_procs), ctx.Queue()
    procs = [ctx.Process(target=_rollout_worker, args=(i, local, n_steps, barrier, q)) for i in range(n_procs)]
    t_spawn = time.perf_counter()
    for p in procs:
        p.start()
    res = []
    try:
        import queue as _queue
        deadline = time.perf_counter() + 180.0
        while len(res) < n_procs:
            try:
                res.append(q.get(timeout=2.0))
            except _queue.Empty:
                if time.perf_counter() > deadline or any(p.exitcode not in (None, 0) for p in procs):
                    raise RuntimeError("a rollout worker process died or timed out (exit codes %s)" % [p.exitcode for p in procs])
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    dts = [r[1] for r in res]
    return {"workload": "configs[3] per-process mode: %d processes x 1 mixture (K=2, sim dtypes: fp16 obs + phases + all-source GT) per env step, "
                        "inputs as host arrays, outputs as host numpy (the sensors' contract)" % n_procs,
            "processes": n_procs, "steps_per_process": n_steps, "env_steps_per_s": n_procs * n_steps / max(dts),
            "latency_ms_per_step_mean": 1e3 * sum(dts) / len(dts) / n_steps, "latency_ms_per_step_worst_process": 1e3 * max(dts) / n_steps,
            "startup_s": time.perf_counter() - t_spawn - max(dts),
            "reference_cpu_ms_per_step": "15-36 ms per env step and core (SURVEY 8a, measured in the survey container)"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    monos, rirs, lens, tc = synth_host_inputs()

    # CPU baseline first (rank 0, N = 1 only), before CUDA is touched: bounded sample of the same workload
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        sample = cpu_sample_size(ref.cores)
        sl = [i % N_DISTINCT for i in range(sample)]
        ref.step(monos[sl], rirs[sl], lens[sl], tc[sl])
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 12.0 and reps < 50):
            cpu_mask = ref.step(monos[sl], rirs[sl], lens[sl], tc[sl])
            reps += 1
        dt = time.perf_counter() - t0
        ref.close()
        cpu_baseline = {"value": sample * reps / dt, "unit": UNIT, "cores": ref.cores, "kind": "port",
                        "sample": f"{reps} x {sample} mixtures of the same seeded workload ({dt:.1f} s): scipy.fftconvolve + restated librosa STFT on a {ref.cores}-process pool, torch-fp32 U-Net on {ref.cores} threads"}
    else:
        cpu_mask, sample = None, 0

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from move2hear_active_av_separation_b200.render import AudioRenderer
    from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
    from oracle import unet as U

    rend = AudioRenderer(dev)
    enc_sd, dec_sd = U.init_state(False, seed=3)
    enc, dec = PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval()
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)

    B = B_PER_GPU
    rep = B // N_DISTINCT
    # each rank renders its own shard: roll the distinct set by rank so shards differ
    roll = (rank * 7) % N_DISTINCT
    h_mono = torch.from_numpy(np.roll(monos, roll, 0)).repeat(rep, 1, 1).pin_memory()
    h_rir = torch.from_numpy(np.roll(rirs, roll, 0)).repeat(rep, 1, 1, 1).pin_memory()
    h_len = torch.from_numpy(np.roll(lens, roll, 0)).repeat(rep, 1).pin_memory()
    h_tc = torch.from_numpy(np.roll(tc, roll, 0)).repeat(rep, 1).pin_memory()
    d_mono, d_rir, d_len, d_tc = (t.to(dev) for t in (h_mono, h_rir, h_len, h_tc))

    stream = torch.cuda.current_stream(dev)

    d_x0 = None if os.environ.get("M2H_BENCH_FP32_OBS") else torch.zeros((B, 18, 18, 128), dtype=torch.bfloat16, device=dev)

    def step_device(timers=None):
        if timers is not None:
            timers[0].record(stream)
        waves = rend.rir_convolve(d_mono, d_rir, d_len)
        if timers is not None:
            timers[1].record(stream)
        # the STFT kernel writes the separator's first-layer operand (bf16, space-to-depth) directly: M2H_BENCH_FP32_OBS=1
        # selects the round-1 chain (fp32 observation in HBM + pack_input_kernel) for comparison
        mix = rend.stft_waves_x0(waves, out=d_x0) if d_x0 is not None else rend.stft_waves(waves)[0]
        if timers is not None:
            timers[2].record(stream)
        with torch.no_grad():
            mask = dec(*enc({"mixed_bin_audio_mag": mix, "target_class": d_tc}))
        if timers is not None:
            timers[3].record(stream)
        return mask

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # parity spot-check of this very configuration against the CPU oracle (not timed)
    mask = step_device()
    if cpu_mask is not None:
        assert roll == 0
        got = mask[[i % N_DISTINCT for i in range(sample)]].float().cpu()
        err = float((got - cpu_mask).norm() / cpu_mask.norm())
        assert err < 2e-2, f"bench parity check failed: mask rel-L2 {err}"

    # clocks are sampled from before the warm-up: nvidia-smi needs ~0.1 s to start streaming and the timed region is ~30 ms
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        time.sleep(0.25)
    t_load0 = time.perf_counter()
    for _ in range(max(args.warmup - 1, 0) + 40):          # W warm-up steps, then ~60 ms more of the same load for the clock samples
        step_device()
    timers = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(timers[i])
    e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1, t_load0) if sampler else None
    ms_total = e0.elapsed_time(e1)
    t_conv = sum(t[0].elapsed_time(t[1]) for t in timers) / args.steps
    t_stft = sum(t[1].elapsed_time(t[2]) for t in timers) / args.steps
    t_unet = sum(t[2].elapsed_time(t[3]) for t in timers) / args.steps
    bytes_stft = BYTES_STFT_X0 if d_x0 is not None else BYTES_STFT

    # ---- end to end: pinned host inputs -> H2D -> render -> U-Net -> D2H of the masks, every step, through the
    # package's host-buffer entry (pipeline.SeparationPipeline: chunks in flight on 3 streams so copies overlap kernels)
    from move2hear_active_av_separation_b200.pipeline import SeparationPipeline
    h_out = torch.empty((B, 2, 512, 32), dtype=torch.float32).pin_memory()
    pipe = SeparationPipeline(enc, dec, dev, K_SRC, N_SAMPLES, L_TAPS, stream_calls=True)      # steps stream through the lanes back to back

    def step_e2e():
        pipe(h_mono, h_rir, h_len, h_tc, h_out)

    for _ in range(2):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, args.steps // 2)
    f0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- the same end-to-end step with the inputs referenced by key: clips and RIRs live in the
    # device-side LRU caches of ingest.CachedAudioRenderer, so a step moves index arrays in and masks out.  The workload
    # tiles N_DISTINCT mixtures, i.e. every key is a hit after the first step - reported next to `e2e`, not instead of it.
    from move2hear_active_av_separation_b200.ingest import CachedAudioRenderer
    cached = CachedAudioRenderer(dev, n_samples=N_SAMPLES, rir_max_taps=L_TAPS, clip_capacity=2 * K_SRC * N_DISTINCT,
                                 rir_capacity=2 * K_SRC * N_DISTINCT, renderer=rend)
    mono_np, rir_np, len_np = h_mono.numpy(), h_rir.numpy(), h_len.numpy()
    clip_keys = [[("clip", (b % N_DISTINCT), k) for k in range(K_SRC)] for b in range(B)]
    rir_keys = [[("rir", (b % N_DISTINCT), k) for k in range(K_SRC)] for b in range(B)]

    def clip_loader(key):
        return mono_np[key[1], key[2]]

    def rir_loader(key):
        return np.ascontiguousarray(rir_np[key[1], key[2], :, :int(len_np[key[1], key[2]])].T)      # wav layout [L, 2]

    side = torch.cuda.Stream(dev)              # masks go back on a second stream: the D2H of step i runs under step i+1

    def step_cached():
        out = cached.render_keys(clip_keys, rir_keys, clip_loader, rir_loader)
        with torch.no_grad():
            mask = dec(*enc({"mixed_bin_audio_mag": out["mix_mag"], "target_class": d_tc}))
        done = torch.cuda.Event()
        done.record(stream)
        with torch.cuda.stream(side):
            side.wait_event(done)
            h_out.copy_(mask.permute(0, 3, 1, 2), non_blocking=True)
            mask.record_stream(side)

    for _ in range(2):
        step_cached()
    stream.wait_stream(side)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for _ in range(e2e_steps):
        step_cached()
    stream.wait_stream(side)
    g1.record(stream)
    barrier()
    ms_cached = g0.elapsed_time(g1)
    cache_stats = cached.stats()

    # ---- BASELINE configs[3]: the rollout step of one rank (14 environments): render + separator inference, eager launches
    # vs one CUDA-graph replay (graphs.GraphedSeparation); device-resident inputs, latency per step
    from move2hear_active_av_separation_b200.graphs import GraphedSeparation
    B_ROLL = 14
    rm, rr, rl, rt = d_mono[:B_ROLL], d_rir[:B_ROLL], d_len[:B_ROLL], d_tc[:B_ROLL]
    graphed = GraphedSeparation(enc, dec, dev, B_ROLL, K_SRC, N_SAMPLES, L_TAPS, renderer=rend)

    def roll_eager():
        waves = rend.rir_convolve(rm, rr, rl)
        mix, _, _ = rend.stft_waves(waves)
        with torch.no_grad():
            return dec(*enc({"mixed_bin_audio_mag": mix, "target_class": rt}))

    def time_calls(fn, n):
        for _ in range(5):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / n

    def latency(fn, n):        # what the RL loop sees: the result is needed before the next environment step
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
            torch.cuda.synchronize(dev)
        return 1e3 * (time.perf_counter() - t0) / n

    ms_roll_eager = time_calls(roll_eager, 50)
    ms_roll_graph = time_calls(lambda: graphed(rm, rr, rl, rt), 50)
    lat_roll_eager = latency(roll_eager, 50)
    lat_roll_graph = latency(lambda: graphed(rm, rr, rl, rt), 50)

    # ---- BASELINE configs[4] (farTarget / unheard stress) as written: K in {3, 4} sources x L in {32000, 48000} taps (2-3 s RIRs:
    # the "taps that reach the window" clipping and the un-pruned first radix stage are exercised), B = 4096 mixtures per GPU,
    # render + binSep inference, every rank independent.  Parity of these shapes: golden cases k3_long / k4_48000.
    # Inputs are generated on the device (6.3 GB of RIRs at the largest point).
    B_ST = 4096
    sweep = []
    st_x0 = None
    gen = torch.Generator(device=dev)
    gen.manual_seed(17 + rank)
    d_tc_st = d_tc.repeat(B_ST // B, 1)
    for K_ST, L_ST in ((3, 32000), (3, 48000), (4, 32000), (4, 48000)):
        st_mono = (3000.0 * torch.randn((B_ST, K_ST, N_SAMPLES), device=dev, generator=gen)).round().clamp(-32767, 32767).to(torch.int16)
        decay = torch.exp(-torch.arange(L_ST, device=dev, dtype=torch.float32) / (L_ST / 6.0))
        st_rir = torch.randn((B_ST, K_ST, 2, L_ST), device=dev, generator=gen)
        st_rir.mul_(decay).mul_(0.03)
        st_len = torch.full((B_ST, K_ST), L_ST, dtype=torch.int32, device=dev)
        st_waves = torch.empty((B_ST, K_ST, 2, N_SAMPLES), dtype=torch.int16, device=dev)
        if d_x0 is not None and st_x0 is None:
            st_x0 = torch.zeros((B_ST, 18, 18, 128), dtype=torch.bfloat16, device=dev)

        def step_stress(tm=None):
            if tm is not None:
                tm[0].record(stream)
            w = rend.rir_convolve(st_mono, st_rir, st_len, out=st_waves)
            if tm is not None:
                tm[1].record(stream)
            mix = rend.stft_waves_x0(w, out=st_x0) if st_x0 is not None else rend.stft_waves(w)[0]      # same hand-over as the headline loop
            if tm is not None:
                tm[2].record(stream)
            with torch.no_grad():
                dec(*enc({"mixed_bin_audio_mag": mix, "target_class": d_tc_st}))
            if tm is not None:
                tm[3].record(stream)

        for _ in range(2):
            step_stress()
        st_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(3)]
        barrier()
        for tm in st_ev:
            step_stress(tm)
        barrier()
        ms_c = sum(t[0].elapsed_time(t[1]) for t in st_ev) / len(st_ev)
        ms_s = sum(t[1].elapsed_time(t[2]) for t in st_ev) / len(st_ev)
        ms_u = sum(t[2].elapsed_time(t[3]) for t in st_ev) / len(st_ev)
        alg = (2 * K_ST * N_SAMPLES + 8 * K_ST * L_ST + 4 * 512 * 32 * 2) * B_ST           # bytes(K, N, L) per mixture
        sweep.append({"K": K_ST, "L": L_ST, "batch_per_gpu": B_ST, "ms_conv": ms_c, "ms_stft": ms_s, "ms_unet": ms_u,
                      "mixtures_per_s": world * B_ST / ((ms_c + ms_s + ms_u) * 1e-3),
                      "render_algorithmic_gb_per_s": alg / ((ms_c + ms_s) * 1e-3) / 1e9,
                      "render_frac_of_hbm_peak": alg / ((ms_c + ms_s) * 1e-3) / 1e9 / peaks()["hbm_gbs"]})
        del st_mono, st_rir, st_len, st_waves
    torch.cuda.empty_cache()

    # ---- batched observation transport - the 14 environments of a rank send (sound ids, RIR ids) requests,
    # the trainer renders them in one batch from the device caches straight into rollout-storage tensors (all 4 audio sensors)
    from move2hear_active_av_separation_b200 import transport as TR
    n_keys = 64
    tr_clips = {f"s{i}": monos[i % N_DISTINCT, i % K_SRC] for i in range(16)}
    tr_rirs = {("scene", 0, r, s_): np.ascontiguousarray(rirs[(r * 2 + s_) % N_DISTINCT, s_ % K_SRC, :, :L_TAPS - 37 * r].T) for r in range(n_keys // 2) for s_ in (0, 1)}
    bor = TR.BatchedObservationRenderer(tr_clips.__getitem__, tr_rirs.__getitem__, device=dev, gt_mono_mag_norm=1.2, rir_capacity=256, clip_capacity=32)
    T_FR = 32
    storage = {"mixed_bin_audio_mag": torch.zeros(21, B_ROLL, 512, T_FR, 2, device=dev), "mixed_bin_audio_phase": torch.zeros(21, B_ROLL, 512, T_FR, 2, device=dev),
               "gt_bin_comps": torch.zeros(21, B_ROLL, 512, T_FR, 4 * K_SRC, device=dev), "gt_mono_comps": torch.zeros(21, B_ROLL, 512, T_FR, 2 * K_SRC, device=dev)}
    rng = np.random.default_rng(5 + rank)

    def make_requests():
        obs = []
        for e in range(B_ROLL):
            req = TR.AudioRenderRequest((f"s{int(rng.integers(16))}", f"s{int(rng.integers(16))}"), "scene", 0, int(rng.integers(n_keys // 2)), (0, 1))
            obs.append({k: TR.DeferredObservation(req, k) for k in TR.AUDIO_SENSORS})
        return obs

    req_steps = [make_requests() for _ in range(20)]
    for i in range(20):                                   # first pass fills the caches (misses)
        bor.insert_into(req_steps[i], storage, 1 + i)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for rep_ in range(3):
        for i in range(20):                               # one episode of MAX_EPISODE_STEPS = 20
            bor.insert_into(req_steps[i], storage, 1 + i)
            torch.cuda.synchronize(dev)                   # the policy needs the observation before the next env step
    dt_tr = (time.perf_counter() - t0) / 60
    transport = {"workload": "SURVEY 8f row 3: 14 envs send (sound ids, RIR ids); one batched render (sim dtypes, phases, all-source GT) from the device "
                             "caches written into RolloutStorage-shaped fp32 tensors, device sync per step",
                 "batch": B_ROLL, "latency_ms_per_step": 1e3 * dt_tr, "env_steps_per_s": world * B_ROLL / dt_tr,
                 "bytes_over_pipe_per_env_step": 4 * 100, "bytes_over_pipe_reference": 2 * 512 * 32 * (2 + 2 + 4 * K_SRC + 2 * K_SRC),
                 "cache": bor.cached.stats()}
    del storage, bor

    # ---- host <-> device copy bandwidth of this rank with every rank copying at the same time: the ceiling of `e2e`
    probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    side2 = torch.cuda.Stream(dev)

    def copy_bw(both):
        barrier()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(4):
            probe_d.copy_(probe_h, non_blocking=True)
            if both:
                with torch.cuda.stream(side2):
                    h_out.view(-1).view(torch.uint8)[:128 << 20].copy_(probe_d[:128 << 20], non_blocking=True)
        stream.wait_stream(side2)
        b2.record(stream)
        barrier()
        return 4 * (256 << 20) / (a.elapsed_time(b2) * 1e-3) / 1e9

    copy_bw(False)
    pcie = {"h2d_gb_per_s_alone": copy_bw(False), "h2d_gb_per_s_with_concurrent_d2h": copy_bw(True), "ranks_copying_concurrently": world}
    del probe_h, probe_d

    # ---- BASELINE configs[2]: separator pretraining step (dataset-path render + both U-Nets fwd/bwd + L1 losses + Adam),
    # B = 256 per GPU, gradients all-reduced (mean) across ranks over NCCL, overlapped with backward
    train = None
    ms_train = 0.0
    if not args.no_train:
        train, ms_train = run_train_step(args, torch, dist, dev, world, rank, rend, d_mono, d_rir, d_len, d_tc, barrier, stream)

    times = torch.tensor([ms_total, ms_e2e, ms_train, ms_cached, -pcie["h2d_gb_per_s_alone"], -pcie["h2d_gb_per_s_with_concurrent_d2h"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_train, ms_cached = (float(times[i]) for i in range(4))
    pcie["h2d_gb_per_s_alone"], pcie["h2d_gb_per_s_with_concurrent_d2h"] = -float(times[4]), -float(times[5])      # slowest rank
    if train is not None:
        train["ms_per_step"] = ms_train / train["steps"]
        train["value"] = world * train["batch_per_gpu"] * train["steps"] / (ms_train * 1e-3)
        train["tflops"] = train["value"] * FLOPS_TRAIN / 1e12
        train["frac_of_bf16_peak"] = train["tflops"] / world / peaks()["bf16_tflops_sustained"]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    rollout_procs = None
    if world == 1 and not args.no_procs:
        del d_mono, d_rir                   # the children allocate their own contexts next to this one
        torch.cuda.empty_cache()
        try:
            rollout_procs = run_rollout_processes(local)
        except Exception as e:           # never lose the line to the process-mode leg
            rollout_procs = {"error": repr(e)[:300]}

    pk = peaks()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (h_mono, h_rir, h_len, h_tc))
    d2h = h_out.numel() * h_out.element_size()
    conv_gbs = BYTES_CONV * B / (t_conv * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("m2h_rir_conv24_kernel", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (render) / bf16 operands + f32 accumulate (U-Net)", "data": "synthetic",
        "config": workload_config({"global_batch": world * B, "parallelism": f"shard-by-mixture x{world}, no data-path collective"}),
        "roofline": {"kernel": "m2h_rir_conv24_kernel", "bound": "hbm", "achieved": conv_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": conv_gbs / pk["hbm_gbs"], "traffic": traffic, "peak_src": pk["src"],
                     "algorithmic_bytes_per_launch": BYTES_CONV * B, "ms_per_launch": t_conv,
                     "note": "an fp32-exact FFT convolution is bound on-chip, not by HBM: ncu (profiles/r2) shows the shared-memory / L1 data pipe 62 % "
                             "and the FP32 pipe 53 % busy, DRAM 15 %; see DESIGN.md section 4.1b"},
        "roofline_other": [
            {"kernel": "m2h_stft_kernel", "bound": "hbm", "achieved": bytes_stft * B / (t_stft * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
             "frac": bytes_stft * B / (t_stft * 1e-3) / 1e9 / pk["hbm_gbs"], "ms_per_launch": t_stft, "algorithmic_bytes_per_mixture": bytes_stft,
             "output": "bf16 first-layer operand of the separator [18,18,128]" if d_x0 is not None else "fp32 observation [512,32,2]"},
            {"kernel": "render chain (conv + stft) vs SURVEY 8d bytes", "bound": "hbm", "achieved": BYTES_RENDER * B / ((t_conv + t_stft) * 1e-3) / 1e9,
             "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": BYTES_RENDER * B / ((t_conv + t_stft) * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": t_conv + t_stft},
            {"kernel": "U-Net forward: 3 unet_slab_kernel + 7 unet_gemm_tma_kernel" + ("" if d_x0 is not None else " + pack_input_kernel"), "bound": "tensor", "achieved": FLOPS_UNET * B / (t_unet * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"],
             "unit": "TFLOP/s", "frac": FLOPS_UNET * B / (t_unet * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], "ms": t_unet},
            # informational: what actually bounds the render kernels on the SM: shared-memory wavefronts per
            # launch are the ncu counts of this workload (profiles/r2/ncu_render_summary.json, B = 1024 per GPU); peak = one
            # 128-byte wavefront per SM and clock at the sampled SM clock.
            {"kernel": "m2h_rir_conv24_kernel (informational)", "bound": "shared-memory data pipe", "achieved": 70.01e6 * B / 1024 / (t_conv * 1e-3) / 1e9,
             "peak": 148 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e-3, "unit": "G wavefronts/s",
             "frac": 70.01e6 * B / 1024 / (t_conv * 1e-3) / 1e9 / (148 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e-3)},
            {"kernel": "m2h_stft_kernel (informational)", "bound": "shared-memory data pipe", "achieved": 50.51e6 * B / 1024 / (t_stft * 1e-3) / 1e9,
             "peak": 148 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e-3, "unit": "G wavefronts/s",
             "frac": 50.51e6 * B / 1024 / (t_stft * 1e-3) / 1e9 / (148 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e-3)},
        ],
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps,
                # what the host link allows at best: every rank uploading its 328 MB per step at the bandwidth the slowest rank
                # reached in `pcie_probe` with all ranks uploading at once (an upper bound: the masks going back cost some of it)
                "pcie_ceiling": world * B / (h2d / (pcie["h2d_gb_per_s_alone"] * 1e9)),
                "frac_of_pcie_ceiling": e2e_value / (world * B / (h2d / (pcie["h2d_gb_per_s_alone"] * 1e9)))},
        "e2e_device_cache": {"value": world * B * e2e_steps / (ms_cached * 1e-3), "unit": UNIT, "ms_per_step": ms_cached / e2e_steps,
                             "h2d_bytes_per_step": 2 * 4 * B * K_SRC, "d2h_bytes_per_step": d2h,
                             "rir_hit_rate": cache_stats["rir_hits"] / max(1, cache_stats["rir_hits"] + cache_stats["rir_misses"]),
                             "note": "same step with clips / RIRs referenced by key and resident in the device-side LRU caches (ingest.py, SURVEY 8f row 1); "
                                     "every key of this tiled workload is a hit after the first step - NOT the headline e2e, which copies every input per step"},
        "rollout_step": {"workload": "configs[3] per-rank rollout step: 14 mixtures (K=2) rendered + binSep U-Net inference, inputs on the device",
                         "batch": B_ROLL, "gpu_ms_eager": ms_roll_eager, "gpu_ms_cuda_graph": ms_roll_graph,
                         "latency_ms_eager": lat_roll_eager, "latency_ms_cuda_graph": lat_roll_graph,
                         "note": "gpu_ms: back-to-back calls timed with CUDA events; latency_ms: wall clock per call with a device synchronise after each (host launch path included)",
                         "mixtures_per_s_cuda_graph": world * B_ROLL / (lat_roll_graph * 1e-3)},
        "rollout_processes": rollout_procs,
        "rollout_transport": transport,
        "stress_sweep": {"workload": "configs[4]: K x L sweep, B = 4096 mixtures per GPU, render + binSep U-Net inference, ranks independent", "points": sweep},
        "pcie_probe": pcie,
        "gpu_launches": (OUR_LAUNCHES_PER_STEP + (0 if d_x0 is not None else 2)) * args.steps,
        "clocks": clocks,
        "train_step": train,
    }
    print(json.dumps(line), flush=True)


B_TRAIN = 256
FLOPS_TRAIN = 3 * (220.7e6 + 201.9e6)        # fwd + dgrad + wgrad of both separators, nominal 2*MACs per mixture


def run_train_step(args, torch, dist, dev, world, rank, rend, d_mono, d_rir, d_len, d_tc, barrier, stream):
    """Times K pretraining steps (PassiveTrainer.train body, passive_trainer.py:218-245) on this rank's shard."""
    from move2hear_active_av_separation_b200.distributed import GradientAllReducer, broadcast_module_state
    from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
    from move2hear_active_av_separation_b200.pretrain import train_step
    torch.manual_seed(1234)
    pol = Move2HearPassiveWoMemoryPolicy().to(dev).train()
    broadcast_module_state(pol)
    opt = torch.optim.Adam(pol.parameters(), lr=5e-4, eps=1e-5, fused=True, capturable=True)       # passive_trainer.py:194-195
    red = None
    if world > 1:
        mode = os.environ.get("M2H_DDP_MODE", "after_backward")       # "overlap": hooked, on a side stream, CTA-limited communicator
        if mode == "overlap":
            from move2hear_active_av_separation_b200.distributed import make_reduction_group
            red = GradientAllReducer(pol.parameters(), modules=pol, overlap=True, group=make_reduction_group(int(os.environ.get("M2H_DDP_CTAS", "8"))))
        else:
            from move2hear_active_av_separation_b200.pretrain import separator_param_groups
            red = GradientAllReducer(pol.parameters(), modules=pol, param_groups=None if mode == "one_group" else separator_param_groups(pol),
                                     start_groups_from_hooks=mode not in ("one_group", "two_groups"))
    B = B_TRAIN
    mono, rir, ln, tc = d_mono[:B], d_rir[:B], d_len[:B], d_tc[:B]

    def step():
        out = rend.render(mono, rir, ln, n_gt=1)                                   # dataset path: fp32 log-mag mixture + target GT magnitude
        mixed, gt_bin = out["mix_mag"], out["gt"][:, 0]
        gt_mono, _ = rend.stft_f32((mono[:, :1].float() * (1.0 / 32768)).contiguous())      # monaural GT magnitude [B,512,T,1]
        b_, m_ = train_step(pol, opt, mixed, gt_bin, gt_mono, tc, grad_sync=red if red else None)
        return b_.detach(), m_.detach()          # keep no autograd graph alive between steps

    def timed(fn, steps):
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record(stream)
        for i in range(steps):
            out = fn()
            marks[i + 1].record(stream)
        barrier()
        per = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
        return out, marks[0].elapsed_time(marks[-1]), per

    losses = []
    for _ in range(3):
        b, m = step()
    losses.append((float(b.detach()), float(m.detach())))
    steps = max(5, args.steps // 2)
    (b, m), ms_eager, per_eager = timed(step, steps)
    # the same step replayed from ONE CUDA graph (pretrain.GraphedTrainStep): the eager step is host-bound (tools/train_hostbound.py)
    mode, ms_total, per_step = "eager", ms_eager, per_eager
    graph_err = None
    if not os.environ.get("M2H_TRAIN_EAGER"):
        try:
            from move2hear_active_av_separation_b200.pretrain import GraphedTrainStep
            gstep = GraphedTrainStep(pol, opt, B, dev, renderer=rend, n_sources=K_SRC, n_samples=N_SAMPLES, rir_taps=L_TAPS,
                                     grad_sync=red if red else None, two_streams=bool(int(os.environ.get("M2H_TRAIN_TWO_STREAMS", "1"))))
            for _ in range(3):
                gstep(mono, rir, ln, target_class=tc)
            (b, m), ms_total, per_step = timed(lambda: gstep(mono, rir, ln, target_class=tc), steps)
            mode = "cuda_graph"
        except Exception as e:       # e.g. a collective that cannot be captured on this NCCL: keep the eager number
            graph_err = repr(e)[:300]
            torch.cuda.synchronize(dev)
    losses.append((float(b.detach()), float(m.detach())))
    info = {"workload": "configs[2] pretraining step: render (dataset path) + binSep & bin2mono U-Nets fwd/bwd (train-mode BN) + L1 losses + fused Adam",
            "batch_per_gpu": B, "steps": steps, "unit": "mixtures/s", "grad_allreduce": ("NCCL AVG in place on the flat gradient arena the wgrad kernels write (32 MB fp32 buckets per module group, no pack / un-pack); "
                               "mode " + os.environ.get("M2H_DDP_MODE", "after_backward") + ": a module group's buckets start from the gradient hook of its last parameter - the separator's under the "
                               "converter's forward and backward, each decoder's under its encoder's backward") if world > 1 else "none (1 GPU)",
            "bin_mono_loss_first_last": losses, "ms_per_step_median_this_rank": per_step[len(per_step) // 2], "ms_per_step_max_this_rank": per_step[-1],
            "mode": mode, "ms_per_step_eager_this_rank": ms_eager / steps, "graph_capture_error": graph_err,
            "graph_note": "one CUDA graph per step; the converter branch and the weight gradients are parallel branches of it (pretrain.train_step side_stream, csrc wgrad fork)"}
    return info, ms_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the pretraining-step measurement (train_step key)")
    ap.add_argument("--no-procs", action="store_true", help="skip the 14-process rollout measurement (rollout_processes key, N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
