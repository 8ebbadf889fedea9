"""Fused Adam for the Robust U-Net training loop.

`rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)` is interchangeable with the reference's
`torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)` (Main_Final.py:552): same hyper-parameters, same
coupled L2 weight decay, same `param_groups` (so `ReduceLROnPlateau`, Main_Final.py:553,622, keeps working), same
`state_dict` layout (`step`, `exp_avg`, `exp_avg_sq`).  One CUDA launch updates all parameter tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, stream_ptr
from .ops import _p

_JOB = np.dtype([("param", "<u8"), ("grad", "<u8"), ("exp_avg", "<u8"), ("exp_avg_sq", "<u8"), ("first_block", "<i8"),
                 ("numel", "<i8")])


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            tab = np.zeros(len(ps), dtype=_JOB)
            first = 0
            step = None
            keep = []                      # contiguous copies of strided gradients: alive until the launch is enqueued
            for i, p in enumerate(ps):
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("rbunet.FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step = st["step"] if step is None else step
                if st["step"] != step:
                    raise RuntimeError("rbunet.FusedAdam: parameters of one group must share the step count")
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                if g.dtype != torch.float32 or g.device != p.device:
                    raise RuntimeError("rbunet.FusedAdam needs fp32 gradients on the parameter's device")
                tab[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), first, p.numel())
                first += (p.numel() + 1023) // 1024
            dev = ps[0].device
            dev_tab = torch.from_numpy(tab.view(np.uint8).copy()).to(dev, non_blocking=True)
            b1, b2 = group["betas"]
            call("rbu_adam_step", _p(dev_tab), len(ps), first, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                 float(group["weight_decay"]), int(step), stream_ptr())
            # the job table and the gradient copies must outlive the asynchronous launch: torch's caching allocator only
            # hands their blocks to later allocations of the SAME stream, which are ordered behind the launch
            self._keep = (dev_tab, keep)
        return loss
