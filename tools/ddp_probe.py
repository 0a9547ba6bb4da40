import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
x = torch.ones(32 << 20, device=dev)
r = timeit(lambda: dist.all_reduce(x))
if rank == 0: print("allreduce 128MB ms (gpu, wall)", r, flush=True)
from move2hear_active_av_separation_b200.distributed import GradientAllReducer, broadcast_module_state
from move2hear_active_av_separation_b200.policy import Move2HearPassiveWoMemoryPolicy
from move2hear_active_av_separation_b200.pretrain import train_step
torch.manual_seed(1)
pol = Move2HearPassiveWoMemoryPolicy().to(dev).train(); broadcast_module_state(pol)
opt = torch.optim.Adam(pol.parameters(), lr=5e-4, eps=1e-5, fused=True)
B = 256
g = torch.Generator(device="cpu").manual_seed(rank)
obs = torch.log1p(torch.rand(B, 512, 32, 2, generator=g) * 3).to(dev); gtb = torch.rand(B, 512, 32, 2, generator=g).to(dev); gtm = torch.rand(B, 512, 32, 1, generator=g).to(dev)
tc = torch.randint(0, 11, (B, 1), generator=g).to(dev)
r = timeit(lambda: train_step(pol, opt, obs, gtb, gtm, tc), 5)
if rank == 0: print("train step, no sync (gpu, wall)", r, flush=True)
red = GradientAllReducer(pol.parameters())
r = timeit(lambda: train_step(pol, opt, obs, gtb, gtm, tc, grad_sync=red.finish), 5)
if rank == 0: print("train step, bucketed all-reduce (gpu, wall)", r, "buckets", len(red.buckets), flush=True)
red.remove()
def flat_sync():
    gs = [p.grad for p in pol.parameters()]
    flat = torch.cat([t.reshape(-1) for t in gs]); dist.all_reduce(flat); flat.div_(2)
    torch._foreach_copy_(gs, [v.view_as(t) for v, t in zip(flat.split([t.numel() for t in gs]), gs)])
r = timeit(lambda: train_step(pol, opt, obs, gtb, gtm, tc, grad_sync=flat_sync), 5)
if rank == 0: print("train step, single flat all-reduce (gpu, wall)", r, flush=True)
dist.destroy_process_group()
