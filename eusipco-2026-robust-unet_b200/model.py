"""Drop-in `torch.nn.Module` surface of the Robust U-Net hot path.

`RobustUNet(n_channels=3, n_classes=1, base_channels=64)` has the constructor, parameter registration
order (hence identical initial weights under the same `torch.manual_seed`), the 290 `state_dict` keys /
shapes / dtypes and the `forward(x) -> probabilities [B,1,H,W]` contract of the reference model
(Main_Final.py:226-321), so reference checkpoints load unchanged and an unmodified
`torch.optim.Adam(model.parameters())` + `loss.backward()` loop (Main_Final.py:552,573-582) trains it.

The sub-modules below are *parameter schemas only*: they own the fp32 master parameters and buffers in
torch layouts, but none of them executes a torch op.  `RobustUNet.forward` hands the whole graph to
`engine.Engine`, which schedules the sm_100a kernels of librbunet.so; the backward pass is one
`torch.autograd.Function` whose gradients come from the same library.  There is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import Engine


class _Slot(nn.Identity):
    """Parameter-free position inside an nn.Sequential (keeps the reference's child indices)."""


def _no_standalone(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter schema; it only runs inside rbunet.RobustUNet "
                       "(the CUDA engine schedules the whole network)")


class ChannelAttention(nn.Module):
    """Parameters of the channel gate (Main_Final.py:82-101): shared MLP C -> C/ratio -> C, no bias."""

    def __init__(self, channels, ratio=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Conv2d(channels, channels // ratio, 1, bias=False), _Slot(),
                                nn.Conv2d(channels // ratio, channels, 1, bias=False))

    forward = _no_standalone


class SpatialAttention(nn.Module):
    """Parameters of the spatial gate (Main_Final.py:104-117): 7x7 conv 2 -> 1, no bias."""

    def __init__(self, kernel_size=7):
        super().__init__()
        if kernel_size != 7:
            raise ValueError("the CUDA spatial-attention kernels are specialised for kernel_size == 7")
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=kernel_size // 2, bias=False)

    forward = _no_standalone


class AttentionGate(nn.Module):
    """Parameters of the skip-connection gate (Main_Final.py:120-148)."""

    def __init__(self, F_g, F_l, F_int):
        super().__init__()
        self.W_g = nn.Sequential(nn.Conv2d(F_g, F_int, 1), nn.BatchNorm2d(F_int))
        self.W_x = nn.Sequential(nn.Conv2d(F_l, F_int, 1), nn.BatchNorm2d(F_int))
        self.psi = nn.Sequential(nn.Conv2d(F_int, 1, 1), nn.BatchNorm2d(1), _Slot())

    forward = _no_standalone


class ResidualBlock(nn.Module):
    """Parameters of one residual block (Main_Final.py:151-196)."""

    def __init__(self, in_channels, out_channels, dropout_rate=0.1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.dropout = nn.Dropout2d(dropout_rate)          # only `.p` is read; masks are drawn by the engine
        self.ca = ChannelAttention(out_channels)
        self.sa = SpatialAttention()
        if in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1, bias=False),
                                          nn.BatchNorm2d(out_channels))
        else:
            self.shortcut = nn.Identity()

    forward = _no_standalone


class DilatedBlock(nn.Module):
    """Parameters of the dilated bottleneck (Main_Final.py:199-223): 1x1 + three 3x3 (d = 1, 2, 4)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        q = out_channels // 4
        self.conv1 = nn.Conv2d(in_channels, q, 1)
        for i, d in ((2, 1), (3, 2), (4, 4)):
            setattr(self, f"conv{i}", nn.Conv2d(in_channels, q, 3, padding=d, dilation=d))
        self.bn = nn.BatchNorm2d(out_channels)

    forward = _no_standalone


class _Graph(torch.autograd.Function):
    """The whole network as one autograd node: forward and backward are kernel schedules of the engine."""

    @staticmethod
    def forward(ctx, model, save, x, *params):
        probs, state = model._engine.forward(x, model.training, save)
        ctx.model = model
        ctx.state = state
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        model, state = ctx.model, ctx.state
        if state is None:
            raise RuntimeError(
                "rbunet.RobustUNet: backward needs a train-mode forward with gradients enabled.  Backward through an "
                "eval-mode forward (running-statistics BatchNorm) is not implemented: the backward kernels apply the "
                "batch-statistics BatchNorm formulas of Main_Final.py:573-582's training loop and would return wrong "
                "gradients, so this raises instead")
        ctx.state = None
        if model._grad_begin_hook is not None:
            model._grad_begin_hook(model._engine)
        grads = model._engine.backward(state, dprobs, allreduce_hook=model._grad_ready_hook)
        if model._grad_transform is not None:
            grads = model._grad_transform(grads)
        out = []
        for name, p in model._named_params():
            g = grads.get(name) if p.requires_grad else None
            out.append(g.reshape(p.shape) if g is not None else None)
        return (None, None, None, *out)


class RobustUNet(nn.Module):
    """B200-native Robust U-Net; same public surface as Main_Final.RobustUNet (Main_Final.py:226-321)."""

    def __init__(self, n_channels=3, n_classes=1, base_channels=64):
        super().__init__()
        if n_classes != 1:
            raise ValueError("the fused sigmoid head supports n_classes == 1 (every reference call site uses 1)")
        b = base_channels
        if b < 16 or b & (b - 1):
            raise ValueError("base_channels must be a power of two >= 16")
        self.inc = ResidualBlock(n_channels, b, 0.1)
        self.down1 = nn.Sequential(_Slot(), ResidualBlock(b, 2 * b, 0.1))
        self.down2 = nn.Sequential(_Slot(), ResidualBlock(2 * b, 4 * b, 0.2))
        self.down3 = nn.Sequential(_Slot(), ResidualBlock(4 * b, 8 * b, 0.2))
        self.bottleneck = nn.Sequential(_Slot(), DilatedBlock(8 * b, 16 * b), ResidualBlock(16 * b, 16 * b, 0.3))
        for k, c in ((4, 8 * b), (3, 4 * b), (2, 2 * b), (1, b)):
            setattr(self, f"att{k}", AttentionGate(c, c, c // 2))
        for k, c, p in ((4, 8 * b, 0.2), (3, 4 * b, 0.2), (2, 2 * b, 0.1), (1, b, 0.1)):
            setattr(self, f"up{k}", nn.ConvTranspose2d(2 * c, c, 2, stride=2))
            setattr(self, f"dec{k}", ResidualBlock(2 * c, c, p))
        self.outc = nn.Sequential(nn.Conv2d(b, n_classes, 1), _Slot())
        self._initialize_weights()
        self._engine = Engine(self)
        self._grad_begin_hook = None      # set by parallel.DataParallel: called with the engine when a backward starts
        self._grad_ready_hook = None      # set by parallel.DataParallel: called with the names of finished grads
        self._grad_transform = None       # set by parallel.DataParallel: swaps in the all-reduced gradients

    def _initialize_weights(self):
        """Kaiming-normal(fan_out, relu) on every Conv2d weight, BN gamma = 1 / beta = 0; ConvTranspose2d and all
        biases keep torch's defaults (Main_Final.py:281-288)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    @property
    def engine(self) -> Engine:
        return self._engine

    # The module tree is fixed after construction, so the (name, Parameter) list is cached: walking 173 parameters through
    # nn.Module.named_parameters() costs ~0.5 ms of host time per call -- time the GPU idles whenever the caller has just
    # synchronised on the previous step (`loss.item()`, Main_Final.py:584).  Anything that can replace Parameter objects
    # goes through _apply / load_state_dict / register_parameter and drops the cache.
    def _named_params(self):
        c = self.__dict__.get("_param_cache")
        if c is None:
            c = tuple(self.named_parameters())
            self.__dict__["_param_cache"] = c
        return c

    def _drop_param_cache(self):
        self.__dict__["_param_cache"] = None

    def _apply(self, fn, recurse=True):
        self._drop_param_cache()
        return super()._apply(fn, recurse)

    def load_state_dict(self, *args, **kwargs):
        self._drop_param_cache()
        return super().load_state_dict(*args, **kwargs)

    def requires_grad_(self, requires_grad: bool = True):
        self._drop_param_cache()
        return super().requires_grad_(requires_grad)

    def forward(self, x):
        params = tuple(p for _, p in self._named_params())
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("rbunet.RobustUNet does not produce a gradient for its input (the first layer's data "
                               "gradient is never computed); detach the input")
        # grad mode is off inside autograd.Function.forward, so decide here whether to keep the backward state.  Only a
        # train-mode forward keeps it: the backward kernels implement batch-statistics BatchNorm (see _Graph.backward)
        save = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _Graph.apply(self, save, x, *params)
