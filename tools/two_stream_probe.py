"""Does running independent (render -> U-Net) chains on two streams raise throughput?  python tools/two_stream_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from move2hear_active_av_separation_b200.render import AudioRenderer
from move2hear_active_av_separation_b200.separator_cnn import PassiveSepDecCNN, PassiveSepEncCNN
from oracle import render as R, unet as U
dev = torch.device("cuda", 0)
B = 1024
monos = torch.from_numpy(np.tile(R.synth_monos(64, 2, seed=0), (16, 1, 1))).to(dev)
rirs, lens = R.synth_rirs(64, 2, L=16000, seed=1)
rir = torch.from_numpy(np.tile(rirs, (16, 1, 1, 1))).to(dev); ln = torch.from_numpy(np.tile(lens, (16, 1))).to(dev)
tc = torch.from_numpy((np.arange(B) % 11).reshape(-1, 1)).to(dev)
enc_sd, dec_sd = U.init_state(False, seed=3)
def chain(n):
    rend = AudioRenderer(dev)
    enc, dec = PassiveSepEncCNN().to(dev).eval(), PassiveSepDecCNN().to(dev).eval()
    enc.load_state_dict(enc_sd); dec.load_state_dict(dec_sd)
    m, r, l, t = monos[:n].contiguous(), rir[:n].contiguous(), ln[:n].contiguous(), tc[:n].contiguous()
    def f():
        w = rend.rir_convolve(m, r, l)
        mix, _, _ = rend.stft_waves(w)
        with torch.no_grad():
            return dec(*enc({"mixed_bin_audio_mag": mix, "target_class": t}))
    return f
def run(fns, streams, iters=20):
    for _ in range(3):
        for f, s in zip(fns, streams):
            with torch.cuda.stream(s): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        for f, s in zip(fns, streams):
            with torch.cuda.stream(s): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters
s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
t = run([chain(1024)], [s0]); print("1 stream  x 1024: %.3f ms  -> %.0f mixtures/s" % (1e3 * t, 1024 / t))
t = run([chain(512), chain(512)], [s0, s1]); print("2 streams x  512: %.3f ms  -> %.0f mixtures/s" % (1e3 * t, 1024 / t))
t = run([chain(1024), chain(1024)], [s0, s1]); print("2 streams x 1024: %.3f ms  -> %.0f mixtures/s" % (1e3 * t, 2048 / t))
s2 = torch.cuda.Stream(dev)
t = run([chain(1024), chain(1024), chain(1024)], [s0, s1, s2]); print("3 streams x 1024: %.3f ms  -> %.0f mixtures/s" % (1e3 * t, 3072 / t))
