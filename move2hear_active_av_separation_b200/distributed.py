"""Data-parallel gradient exchange for separator pretraining (BASELINE config 3; SURVEY 8e: the path's one collective).

The reference pretrains on one GPU; data-parallel pretraining follows ``DistributedDataParallel`` semantics:
every rank runs its own shard (per-rank BatchNorm statistics), parameter gradients are averaged across ranks once per
step, running statistics follow rank 0.  ``GradientAllReducer`` packs the gradients into a few flat fp32 buckets in
reverse parameter order (decoder first: the order backward produces them) and all-reduces them (NCCL ``AVG``) between
``backward()`` and ``optimizer.step()``; optionally a bucket is reduced on a side stream as soon as its last gradient has
been accumulated.
Works with any ``torch.distributed`` backend (``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradientAllReducer:
    """``overlap=False`` (default): :meth:`finish` all-reduces the buckets on the current stream after backward - 0.36 ms for
    the 134 MB of both separators on 4 B200s (NVLink 5 / NVSwitch), plus one pack and one fused un-pack kernel per bucket.
    ``overlap=True``: buckets are reduced on a side stream from gradient hooks while backward still runs.  Measured on
    B200 this is SLOWER (+3.3 ms per step on 4 GPUs against +0.9 ms): the backward kernels are persistent and fill every SM, so
    the NCCL kernels of the side stream are scheduled piecemeal and every rank waits for the slowest one."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None, overlap: bool = False,
                 modules: torch.nn.Module = None, param_groups=None, start_groups_from_hooks: bool = False):
        """``modules`` (optional): a module tree containing the separator modules of ``separator_cnn``.  Their backward then
        writes every parameter gradient straight into this reducer's flat per-bucket buffers (the autograd Functions hand
        out views of the arena, which ``AccumulateGrad`` adopts as ``p.grad`` without a copy), so a bucket is all-reduced
        IN PLACE: no ``torch.cat`` pack, no un-pack copy (each cost a 134 MB read + write per step)."""
        """``param_groups`` (optional list of parameter lists, together = ``params``): buckets never straddle two groups and
        :meth:`start` reduces one group early - ``pretrain.train_step`` starts the binaural separator's group as soon as its
        backward is done, under the converter's backward (the two passes are independent: detached masks).
        ``start_groups_from_hooks``: a group is started by the gradient hook of its last parameter, i.e. in the middle of a
        backward pass as soon as the autograd Function that owns the group has returned (decoder gradients travel while the
        encoder's backward runs)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.overlap = overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # buckets in reverse registration order (gradients of late layers are ready first)
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._bucket_group: List[int] = []
        if param_groups is None:
            param_groups = [self.params]
        else:
            param_groups = [[p for p in g if p.requires_grad] for g in param_groups]
            if sorted(id(p) for g in param_groups for p in g) != sorted(id(p) for p in self.params):
                raise ValueError("param_groups must partition params")
        for gi in reversed(range(len(param_groups))):
            cur, cur_bytes = [], 0
            for p in reversed(param_groups[gi]):
                cur.append(p)
                cur_bytes += p.numel() * 4
                if cur_bytes >= bucket_bytes:
                    self.buckets.append(cur)
                    self._bucket_group.append(gi)
                    cur, cur_bytes = [], 0
            if cur:
                self.buckets.append(cur)
                self._bucket_group.append(gi)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._arena = None            # per bucket: flat fp32 buffer; _slot[id(p)] = (bucket, offset)
        self._slot = {}
        if modules is not None:
            self._arena = []
            for i, b in enumerate(self.buckets):
                o = 0
                for p in b:
                    self._slot[id(p)] = (i, o)
                    o += (p.numel() + 3) & ~3                       # 16-byte aligned slots
                self._arena.append(torch.zeros(o, dtype=torch.float32, device=b[0].device))
            for m in modules.modules():
                if hasattr(m, "_grad_struct"):
                    m._grad_arena = self
        self._pending = [0] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._stream = None
        self._hooks = []
        self._group_of = {id(p): gi for gi, g in enumerate(param_groups) for p in g}
        self._group_size = [len(g) for g in param_groups]
        self._group_pending = list(self._group_size)
        if self.world > 1 and self.overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        elif self.world > 1 and start_groups_from_hooks:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_group_grad))
        self._reset()

    def _on_group_grad(self, p):
        gi = self._group_of[id(p)]
        self._group_pending[gi] -= 1
        if self._group_pending[gi] == 0:
            self.start(gi)

    def _reset(self):
        self._group_pending = list(self._group_size)
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _avg_op(self):
        """NCCL averages inside the collective; other backends (gloo in the CPU tests) sum and :meth:`finish` divides."""
        return dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM

    def grad_view(self, p: torch.nn.Parameter) -> torch.Tensor:
        """A FRESH view (so that autograd may adopt it) of the arena slot that receives ``p``'s gradient."""
        i, o = self._slot[id(p)]
        return self._arena[i][o:o + p.numel()].view(p.shape)

    def _in_arena(self, i: int) -> bool:
        if self._arena is None:
            return False
        base = self._arena[i].data_ptr()
        return all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                   and p.grad.data_ptr() == base + 4 * self._slot[id(p)][1] for p in self.buckets[i])

    def _launch(self, i: int):
        bucket = self.buckets[i]
        if self._in_arena(i):
            flat = self._arena[i]                       # the gradients already live in the flat buffer: reduce in place
        else:
            flat = torch.cat([p.grad.reshape(-1).float() for p in bucket])
        self._flat[i] = flat
        if not self.overlap:
            self._work[i] = dist.all_reduce(flat, op=self._avg_op(), group=self.group, async_op=True)
        elif flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(flat.device)
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                self._work[i] = dist.all_reduce(flat, op=self._avg_op(), group=self.group, async_op=True)
        else:
            self._work[i] = dist.all_reduce(flat, op=self._avg_op(), group=self.group, async_op=True)

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def start(self, group_index: int):
        """Launch the all-reduce of every bucket of one parameter group NOW (its gradients are complete on the current stream);
        :meth:`finish` launches the rest and waits for all of them."""
        if self.world <= 1 or self.overlap:
            return
        for i, bucket in enumerate(self.buckets):
            if self._bucket_group[i] == group_index and self._work[i] is None and all(p.grad is not None for p in bucket):
                self._launch(i)

    def __call__(self):
        self.finish()

    def finish(self):
        """Wait for every bucket, write the averaged gradients back.  Call between ``backward()`` and ``optimizer.step()``."""
        if self.world <= 1:
            return
        for i, bucket in enumerate(self.buckets):
            if self._work[i] is None:          # some gradient of the bucket never arrived (unused parameter): reduce what exists
                for p in bucket:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(i)
            self._work[i].wait()
            flat = self._flat[i]
            if flat.is_cuda and self._stream is not None:
                torch.cuda.current_stream(flat.device).wait_stream(self._stream)
            if self._avg_op() != dist.ReduceOp.AVG:
                flat.div_(self.world)
            if self._arena is not None and flat is self._arena[i]:
                continue                                # reduced in place
            views, o = [], 0
            for p in bucket:
                n = p.numel()
                views.append(flat[o:o + n].view_as(p.grad))
                o += n
            torch._foreach_copy_([p.grad for p in bucket], views)          # one fused kernel per bucket, not one copy per parameter
        self._reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def make_reduction_group(max_ctas: int = 8):
    """A dedicated NCCL communicator for the gradient all-reduce whose kernels use at most ``max_ctas`` CTAs
    (``ncclConfig_t.maxCTAs``).  The backward kernels of this path are persistent and fill every SM; an all-reduce launched
    next to them with NCCL's default CTA count is scheduled piecemeal and every rank waits for the slowest (measured in round 1:
    the hooked overlap was slower than reducing after backward).  A small CTA budget lets the collective run in the gaps:
    NVLink 5 needs only a few CTAs to saturate (NVLS reduces inside the switch)."""
    if not dist.is_initialized() or dist.get_backend() != "nccl":
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas)
    opts.config.min_ctas = 1
    return dist.new_group(backend="nccl", pg_options=opts)


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None):
    """Parameters and buffers (BatchNorm running statistics) follow rank ``src``, as DDP does at construction / each forward."""
    if not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
