"""Scratch: quick device timing of the two render kernels (superseded by bench.py)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from move2hear_active_av_separation_b200.render import AudioRenderer
from oracle import render as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = 2
rend = AudioRenderer(torch.device('cuda', 0))
monos = torch.from_numpy(R.synth_monos(8, K, seed=0)).cuda().repeat(B // 8, 1, 1)
rirs_np, lens_np = R.synth_rirs(8, K, seed=1)
rirs = torch.from_numpy(rirs_np).cuda().repeat(B // 8, 1, 1, 1)
lens = torch.from_numpy(lens_np).cuda().repeat(B // 8, 1)
for it in range(3):
    out = rend.render(monos, rirs, lens)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
n = 10
ev[0].record()
for it in range(n):
    waves = rend.rir_convolve(monos, rirs, lens)
ev[1].record()
for it in range(n):
    rend.stft_waves(waves)
ev[2].record()
torch.cuda.synchronize()
tc = ev[0].elapsed_time(ev[1]) / n
ts = ev[1].elapsed_time(ev[2]) / n
print(f"B={B} conv {tc:.3f} ms  stft {ts:.3f} ms  total {tc+ts:.3f} ms -> {B/(tc+ts)*1e3:.0f} mixtures/s; HBM-alg {451072*B/(tc+ts)/1e6:.1f} GB/s")
