"""U-Net forward alone at the bench batch: python tools/unet_fwd_probe.py [B] [iters]   (also the ncu target of tools/unet_iter.sh)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
from oracle import unet as U
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
enc_sd, dec_sd = U.init_state(False, seed=3)
enc, dec = PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval()
enc.load_state_dict(enc_sd); dec.load_state_dict(dec_sd)
g = torch.Generator().manual_seed(0)
obs = torch.log1p(torch.rand(B, 512, 32, 2, generator=g) * 4).to(dev)
tc = torch.from_numpy((np.arange(B) % 11).reshape(-1, 1)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def f():
    with torch.no_grad():
        return dec(*enc({"mixed_bin_audio_mag": obs, "target_class": tc}))
for _ in range(3): f()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("B=%d U-Net forward: median %.4f ms  min %.4f ms (L2 flushed between iterations)" % (B, ts[len(ts) // 2], ts[0]))
