"""Batch-sharded data parallelism for the Robust U-Net hot path (one process per GPU).

The reference is single-process; north_star asks for a data-parallel training step with a bucketed
gradient all-reduce over NVLink overlapped with the backward pass.  Images are independent in every kernel of the
path except the train-mode BatchNorm statistics, which stay per replica (plain-DDP semantics; the reference has no
SyncBN), so the only exchange per step is the gradient all-reduce:

  * parameters are grouped into flat fp32 buckets in the order the engine finishes their gradients
    (outc, dec1/att1/up1, ..., dec4/att4/up4, bottleneck, down3, down2, down1, inc);
  * the weight-gradient GEMMs write straight into their bucket slices (`Engine.grad_alloc`); when the engine reports a
    stage finished (`Engine.backward(..., allreduce_hook=...)`) only the small per-channel gradients are copied (one
    multi-tensor launch per stage).  A bucket whose members are all present is all-reduced (average) on a dedicated
    communication stream, which waits for the compute stream AND for the weight-gradient side stream through events --
    the compute stream itself never waits and carries on with the next stage's kernels;
  * at the end of the backward the compute stream waits for the communication stream and the parameters receive
    views of the reduced buckets as `.grad`.

`GradBucketer` is device-agnostic (CPU tensors + gloo in the tests, CUDA tensors + NCCL on the box).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

STAGE_ORDER = ("outc", "dec1", "att1", "up1", "dec2", "att2", "up2", "dec3", "att3", "up3", "dec4", "att4", "up4",
               "bottleneck", "down3", "down2", "down1", "inc")


def reverse_execution_order(names: Iterable[str]) -> List[str]:
    """Parameter names sorted by the stage in which Engine.backward completes their gradient."""
    rank = {s: i for i, s in enumerate(STAGE_ORDER)}
    names = list(names)
    return sorted(names, key=lambda n: (rank[n.split(".")[0]], names.index(n)))


class GradBucketer:
    def __init__(self, named_shapes: Sequence[Tuple[str, torch.Size]], device, process_group=None,
                 bucket_bytes: int = 25 << 20):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.device = torch.device(device)
        order = reverse_execution_order([n for n, _ in named_shapes])
        shapes = dict(named_shapes)
        self.slots: Dict[str, Tuple[int, int, int]] = {}     # name -> (bucket, offset, numel)
        self.bucket_members: List[List[str]] = []
        sizes: List[int] = []
        cur, cur_elems = [], 0
        for n in order:
            numel = int(torch.Size(shapes[n]).numel())
            if cur and (cur_elems + numel) * 4 > bucket_bytes:
                self.bucket_members.append(cur)
                sizes.append(cur_elems)
                cur, cur_elems = [], 0
            self.slots[n] = (len(self.bucket_members), cur_elems, numel)
            cur.append(n)
            cur_elems += (numel + 3) // 4 * 4          # keep every slice 16-byte aligned
        if cur:
            self.bucket_members.append(cur)
            sizes.append(cur_elems)
        self.shapes = shapes
        self.flat = [torch.zeros(s, dtype=torch.float32, device=self.device) for s in sizes]
        self.cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._reset()

    def _reset(self):
        self.pending = [set(m) for m in self.bucket_members]
        self.works = []
        self.launched = [False] * len(self.flat)
        self.side_events = [None] * len(self.flat)

    def slot_view(self, name: str, shape=None):
        """The bucket slice that holds `name`'s gradient, shaped like the parameter (the engine's weight-gradient kernels
        write into it directly); None for a parameter outside the buckets."""
        ent = self.slots.get(name)
        if ent is None:
            return None
        b, off, numel = ent
        shp = self.shapes[name] if shape is None else shape
        if int(torch.Size(shp).numel()) != numel:
            return None
        return self.flat[b][off:off + numel].view(shp)

    # -- called by the engine hook -------------------------------------------------------------
    def ready(self, names: Iterable[str], grads: Dict[str, torch.Tensor], side_event=None):
        """`grads[n]` is complete for every n in names (on the current stream, or -- for tensors written by kernels of
        another stream -- behind `side_event`).  Gradients that already live in their bucket slice are not copied."""
        touched = set()
        dsts, srcs = [], []
        for n in names:
            if n not in self.slots:
                continue
            b, off, numel = self.slots[n]
            g = grads[n]
            dst = self.flat[b][off:off + numel]
            if g.data_ptr() != dst.data_ptr():
                dsts.append(dst)
                srcs.append(g.reshape(-1))
            self.pending[b].discard(n)
            touched.add(b)
        if dsts:
            if self.cuda:
                torch._foreach_copy_(dsts, srcs)           # one multi-tensor launch for the stage's small gradients
            else:
                for d, s_ in zip(dsts, srcs):
                    d.copy_(s_)
        for b in sorted(touched):
            if side_event is not None:
                self.side_events[b] = side_event            # the side stream is in-order: the latest event covers the rest
            if not self.pending[b] and not self.launched[b]:
                self._launch(b)

    def _launch(self, b: int):
        self.launched[b] = True
        if self.world == 1:
            return
        buf = self.flat[b]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if self.side_events[b] is not None:
                    self.comm_stream.wait_event(self.side_events[b])
                self.works.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        else:  # gloo has no AVG
            self.works.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), buf))

    # -- called once at the end of the backward -------------------------------------------------
    def finish(self) -> Dict[str, torch.Tensor]:
        for b in range(len(self.flat)):
            if not self.launched[b]:                   # parameters without a gradient this step: reduce what is there
                self._launch(b)
        for w in self.works:
            if self.cuda:
                w.wait()                               # makes the current stream wait for the NCCL stream
            else:
                work, buf = w
                work.wait()
                buf.div_(self.world)
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        out = {n: self.flat[b][off:off + numel].view(self.shapes[n]) for n, (b, off, numel) in self.slots.items()}
        self._reset()
        return out


class DataParallel(nn.Module):
    """Wraps rbunet.RobustUNet: same forward, gradients averaged over the process group during backward."""

    def __init__(self, module: nn.Module, process_group=None, bucket_bytes: int = 25 << 20, broadcast_buffers: bool = False):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("rbunet.DataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.pg = process_group
        self.broadcast_buffers = broadcast_buffers
        dev = next(module.parameters()).device
        for t in list(module.parameters()) + list(module.buffers()):        # replicas start identical (rank 0 wins)
            dist.broadcast(t.data, src=0, group=process_group)
        self.bucketer = GradBucketer([(n, p.shape) for n, p in module.named_parameters() if p.requires_grad], dev,
                                     process_group, bucket_bytes)
        self._params = [p for p in module.parameters() if p.requires_grad]
        module._grad_begin_hook = self._begin
        module._grad_ready_hook = self.bucketer.ready
        module._grad_transform = self._finish

    def _detach_surviving_grads(self):
        """The gradients handed to autograd are views of the flat buckets, which the next backward overwrites.  A .grad
        that survived (zero_grad(set_to_none=False), gradient accumulation) would be overwritten and then accumulated
        into itself (2x): give such a .grad its own storage before the buckets are touched."""
        lo_hi = [(b.data_ptr(), b.data_ptr() + b.numel() * 4) for b in self.bucketer.flat]
        for p in self._params:
            g = p.grad
            if g is not None and any(lo <= g.data_ptr() < hi for lo, hi in lo_hi):
                p.grad = g.clone()

    def _begin(self, engine):
        """Start of a backward pass: surviving .grad views leave the buckets, then the engine's weight-gradient kernels
        are pointed at the bucket slices."""
        self._detach_surviving_grads()
        engine.grad_alloc = self.bucketer.slot_view

    def _finish(self, grads):
        self.module._engine.grad_alloc = None
        return self.bucketer.finish()

    def forward(self, x):
        if self.broadcast_buffers and self.training:
            for b in self.module.buffers():
                dist.broadcast(b.data, src=0, group=self.pg)
        return self.module(x)
