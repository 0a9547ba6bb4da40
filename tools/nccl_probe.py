import os, torch, torch.distributed as dist, time
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(32 << 20, device="cuda")  # 128 MB fp32
for _ in range(3): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
if dist.get_rank() == 0: print("allreduce 128MB ms", e0.elapsed_time(e1) / 10)
# async on side stream as GradientAllReducer does
s = torch.cuda.Stream()
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        w = dist.all_reduce(x, async_op=True)
    w.wait(); torch.cuda.current_stream().wait_stream(s)
e1.record(); torch.cuda.synchronize()
if dist.get_rank() == 0: print("async allreduce 128MB ms", e0.elapsed_time(e1) / 10)
dist.destroy_process_group()
