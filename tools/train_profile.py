"""Launch list of ONE pretraining step (B = 256): python tools/train_profile.py (run under ncu --metrics gpu__time_duration.sum)."""
import os
import sys

import numpy as np
import torch

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
rir = torch.from_numpy(np.tile(rirs, (4, 1, 1, 1))).to(dev)
ln = torch.from_numpy(np.tile(lens, (4, 1))).to(dev)
tc = torch.from_numpy(np.random.default_rng(0).integers(0, 102, (B, 1))).to(dev)
torch.manual_seed(1234)
pol = Move2HearPassiveWoMemoryPolicy().to(dev).train()
opt = torch.optim.Adam(pol.parameters(), lr=5e-4, eps=1e-5, fused=True)


def step():
    out = rend.render(mono, rir, ln, n_gt=1)
    gt_mono, _ = rend.stft_f32((mono[:, :1].float() * (1.0 / 32768)).contiguous())
    return train_step(pol, opt, out["mix_mag"], out["gt"][:, 0], gt_mono, tc)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
