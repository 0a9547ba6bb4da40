"""Is the pretraining step bound by the host (Python + launches) or by the GPU?  python tools/train_hostbound.py
Prints, per step: host enqueue time (no synchronisation inside the loop), GPU time (events), and the same with the GPU kept busy."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
from move2hear_active_av_separation_b200.pretrain import train_step
from move2hear_active_av_separation_b200.render import AudioRenderer
from oracle import render as R
dev = torch.device("cuda", 0)
B = 256
rend = AudioRenderer(dev)
mono = torch.from_numpy(np.tile(R.synth_monos(64, 2, seed=0), (4, 1, 1))).to(dev)
rirs, lens = R.synth_rirs(64, 2, L=16000, seed=1)
rir = torch.from_numpy(np.tile(rirs, (4, 1, 1, 1))).to(dev); ln = torch.from_numpy(np.tile(lens, (4, 1))).to(dev)
tc = torch.from_numpy(np.random.default_rng(0).integers(0, 11, (B, 1))).to(dev)
torch.manual_seed(1234)
pol = Move2HearPassiveWoMemoryPolicy().to(dev).train()
opt = torch.optim.Adam(pol.parameters(), lr=5e-4, eps=1e-5, fused=True)
def step():
    out = rend.render(mono, rir, ln, n_gt=1)
    gt_mono, _ = rend.stft_f32((mono[:, :1].float() * (1.0 / 32768)).contiguous())
    return train_step(pol, opt, out["mix_mag"], out["gt"][:, 0], gt_mono, tc)
for _ in range(5): step()
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): step()
e1.record(); t_enq = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print("host enqueue %.3f ms/step   gpu (events) %.3f ms/step   wall incl. drain %.3f ms/step" % (1e3 * t_enq / n, e0.elapsed_time(e1) / n, 1e3 * t_all / n))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("single step from an empty queue: host enqueue %.3f ms, until the GPU drained %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
# phases on the host
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
