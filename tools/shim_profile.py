"""Where does one env step of the B = 1 drop-in shim spend its time?  python tools/shim_profile.py   (single process)"""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from move2hear_active_av_separation_b200 import render as MR
from oracle import render as R
monos = R.synth_monos(1, 2, 16000, seed=1000)[0]
rp, ln = R.synth_rirs(8, 2, 16000, seed=2000, ragged=True)
rirs = [[R.rir_wav_layout(rp[j, k], int(ln[j, k])) for k in range(2)] for j in range(8)]
sim = SimpleNamespace(_cached_mono_waveforms={}, _cached_mono_specs={}, config=SimpleNamespace(AUDIO=SimpleNamespace(NORM_TYPE="l2")))
names = ["a", "b"]
for j in range(5):
    MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)
n = 300
t0 = time.perf_counter()
for j in range(n):
    MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)
dt = time.perf_counter() - t0
print("single process: %.3f ms per env step" % (1e3 * dt / n))
# GPU time of the same step
rend = MR.get_renderer(torch.device("cuda", 0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for j in range(50):
    MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)
e1.record(); torch.cuda.synchronize()
print("device span per step (includes host gaps): %.3f ms" % (e0.elapsed_time(e1) / 50))
pr = cProfile.Profile(); pr.enable()
for j in range(n):
    MR.render_sim_arrays(sim, list(monos), rirs[j % 8], names, 1.2)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
