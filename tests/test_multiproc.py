"""CPU, world_size 2, gloo: the host-side sharding / max-over-ranks logic bench.py uses for N > 1."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from oracle import render as R
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N_DISTINCT = 8
    monos = R.synth_monos(N_DISTINCT, 2, seed=0)
    # same sharding rule as bench.py: every rank rolls the seeded set by 7*rank
    roll = (rank * 7) %% N_DISTINCT
    mine = np.roll(monos, roll, 0)
    digest = torch.tensor([float(np.abs(mine[0]).sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    t = torch.tensor([10.0 + rank, 3.0 - rank], dtype=torch.float64)      # per-rank (value, e2e) step times
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "digests": [float(g) for g in gathered], "max": t.tolist()}))
    dist.destroy_process_group()
""") % ROOT


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_sharding_and_max_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["max"] == [11.0, 3.0]
    assert d["digests"][0] != d["digests"][1]          # the two ranks render different shards


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


GRAD_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from move2hear_active_av_separation_b200.distributed import GradientAllReducer, broadcast_module_state
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(100 + rank)                       # ranks start from different weights ...
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    broadcast_module_state(net)                         # ... and follow rank 0 afterwards
    red = GradientAllReducer(net.parameters(), bucket_bytes=256, overlap=os.environ.get("M2H_TEST_OVERLAP") == "1")      # tiny buckets: several per step
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(12, 8, generator=g), torch.randn(12, 4, generator=g)
    x, y = x_all[rank * 6:(rank + 1) * 6], y_all[rank * 6:(rank + 1) * 6]
    loss = torch.nn.functional.mse_loss(net(x), y)
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.finish()
    # expected: mean over ranks of the local gradients
    ok = True
    for p, l in zip(net.parameters(), local):
        gathered = [torch.zeros_like(l) for _ in range(world)]
        dist.all_gather(gathered, l)
        ok = ok and torch.allclose(p.grad, sum(gathered) / world, atol=1e-6)
    w0 = [torch.zeros_like(net[0].weight) for _ in range(world)]
    dist.all_gather(w0, net[0].weight.data)
    same_init = bool(torch.equal(w0[0], w0[1]))
    # arena mode: a module whose backward hands out views of the reducer's flat buffers (what separator_cnn does on CUDA)
    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, mod, x, w):
            ctx.mod, ctx.x = mod, x
            return x @ w
        @staticmethod
        def backward(ctx, dy):
            g = ctx.mod._grad_arena.grad_view(ctx.mod.w)
            g.copy_(ctx.x.t() @ dy)
            return None, None, g
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(8, 4))
        def _grad_struct(self, device=None):
            pass
        def forward(self, x):
            return Fn.apply(self, x, self.w)
    toy = Toy()
    broadcast_module_state(toy)
    red2 = GradientAllReducer(toy.parameters(), modules=toy)
    torch.nn.functional.mse_loss(toy(x), y).backward()
    in_place = toy.w.grad.data_ptr() == red2._arena[0].data_ptr() and red2._in_arena(0)
    local = toy.w.grad.clone()
    red2.finish()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    arena_ok = bool(in_place and torch.allclose(toy.w.grad, sum(gathered) / world, atol=1e-6) and toy.w.grad.data_ptr() == red2._arena[0].data_ptr())
    # parameter groups: two independent halves, the first one reduced early (start) while the second is still being computed
    a, b = torch.nn.Linear(8, 4), torch.nn.Linear(8, 4)
    both = torch.nn.ModuleList([a, b])
    broadcast_module_state(both)
    red3 = GradientAllReducer(both.parameters(), bucket_bytes=64, param_groups=[list(a.parameters()), list(b.parameters())])
    groups_split = all(len({id(q) for q in bk} & {id(q) for q in a.parameters()}) in (0, len(bk)) for bk in red3.buckets)
    torch.nn.functional.mse_loss(a(x), y).backward()
    la = [q.grad.clone() for q in a.parameters()]
    red3.start(0)
    started = [red3._work[i] is not None for i in range(len(red3.buckets))]
    torch.nn.functional.mse_loss(b(x), y).backward()
    lb = [q.grad.clone() for q in b.parameters()]
    red3()
    groups_ok = groups_split and started == [g == 0 for g in red3._bucket_group]
    for q, l in zip(list(a.parameters()) + list(b.parameters()), la + lb):
        gathered = [torch.zeros_like(l) for _ in range(world)]
        dist.all_gather(gathered, l)
        groups_ok = groups_ok and torch.allclose(q.grad, sum(gathered) / world, atol=1e-6)
    # the same with hook-started groups: no explicit start()
    for q in both.parameters():
        q.grad = None
    red4 = GradientAllReducer(both.parameters(), bucket_bytes=64, param_groups=[list(a.parameters()), list(b.parameters())], start_groups_from_hooks=True)
    torch.nn.functional.mse_loss(a(x), y).backward()
    hooks_ok = [red4._work[i] is not None for i in range(len(red4.buckets))] == [g == 0 for g in red4._bucket_group]
    torch.nn.functional.mse_loss(b(x), y).backward()
    hooks_ok = hooks_ok and all(w is not None for w in red4._work)
    red4()
    for q, l in zip(list(a.parameters()) + list(b.parameters()), la + lb):
        gathered = [torch.zeros_like(l) for _ in range(world)]
        dist.all_gather(gathered, l)
        hooks_ok = hooks_ok and torch.allclose(q.grad, sum(gathered) / world, atol=1e-6)
    groups_ok = groups_ok and hooks_ok
    if rank == 0:
        print(json.dumps({"ok": bool(ok), "same_init": same_init, "n_buckets": len(red.buckets), "arena_ok": arena_ok, "groups_ok": bool(groups_ok)}))
    dist.destroy_process_group()
""") % ROOT


@pytest.mark.parametrize("overlap", ["0", "1"], ids=["after_backward", "hooked_overlap"])
def test_two_rank_gloo_gradient_allreduce(tmp_path, overlap):
    script = tmp_path / "grad_worker.py"
    script.write_text(GRAD_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)],
                         capture_output=True, text=True, timeout=240, env=dict(os.environ, M2H_TEST_OVERLAP=overlap))
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["ok"] and d["same_init"] and d["n_buckets"] > 1
    assert d["arena_ok"]            # gradients written into the flat arena are adopted by autograd and reduced in place
    assert d["groups_ok"]           # buckets respect parameter groups; start(0) launches only the first group's
