"""GPU parity: tcgen05 weight-gradient GEMM (rbu_wgrad_gemm through the C ABI) against torch CPU autograd on the
same bf16-rounded operands.  fp32 accumulation, fp32 output: tolerance is accumulation-order noise only."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import bf16r, rel_l2, to_view

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


CASES = [
    # name, N, H, W, Cin, Cout, ksz, dil
    ("3x3_64_64", 2, 16, 16, 64, 64, 3, 1),
    ("3x3_128_256", 2, 16, 16, 128, 256, 3, 1),
    ("3x3_d2", 1, 16, 16, 64, 128, 3, 2),
    ("3x3_d4", 2, 16, 16, 128, 64, 3, 4),
    ("1x1_128_64", 2, 16, 16, 128, 64, 1, 1),
    ("1x1_32_16", 2, 8, 8, 32, 16, 1, 1),
    ("3x3_16_32_small", 2, 8, 8, 16, 32, 3, 1),
    ("3x3_ragged", 3, 24, 12, 64, 96, 3, 1),
    ("3x3_tiny_spatial", 5, 2, 2, 128, 64, 3, 1),
    ("3x3_many_tiles", 4, 64, 64, 64, 64, 3, 1),
    ("1x1_stem_like", 2, 32, 32, 32, 128, 1, 1),
    ("3x3_512_512", 2, 8, 8, 512, 512, 3, 1),
    # second form of the halo kernel (128 output x 32 input channels per item, three taps stacked in N = 96): ragged tile
    # edges, Cin = 32 / 96 (one and three 32-channel blocks), several output blocks, many tiles per item
    ("3x3_form2_ragged", 3, 24, 40, 96, 128, 3, 1),
    ("3x3_form2_cin32", 2, 16, 32, 32, 256, 3, 1),
    ("3x3_form2_many_tiles", 4, 64, 64, 64, 128, 3, 1),
    ("3x3_form2_min_tile", 2, 8, 16, 128, 128, 3, 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_wgrad(case):
    from rbunet.engine import Engine
    name, N, H, W, Cin, Cout, ksz, dil = case
    dev = torch.device("cuda:0")
    x = bf16r(_rand((N, Cin, H, W), 1))
    dy = bf16r(_rand((N, Cout, H, W), 2))
    w = torch.zeros((Cout, Cin, ksz, ksz), requires_grad=True)
    F.conv2d(x, w, padding=dil * (ksz // 2), dilation=dil).backward(dy)
    eng = Engine(None)
    out = torch.full((Cout, Cin, ksz, ksz), 7.0, device=dev)
    # operands as channel slices of wider buffers to exercise ld != C
    eng.wgrad(N, H, W, to_view(dy, dev, ld=Cout + 16, off=8, fill=3.0), to_view(x, dev, ld=Cin + 8, off=8, fill=5.0),
              ksz * ksz, dil, False, out)
    torch.cuda.synchronize()
    err = rel_l2(out, w.grad)
    assert err < 1e-4, f"{name}: rel-L2 {err}"


def test_conv_transpose_wgrad():
    from rbunet.engine import Engine
    dev = torch.device("cuda:0")
    for (N, H, W, Cin, Cout) in ((2, 8, 8, 128, 64), (1, 4, 4, 256, 128), (3, 6, 10, 64, 32)):
        x = bf16r(_rand((N, Cin, H, W), 3))
        dout = bf16r(_rand((N, Cout, 2 * H, 2 * W), 4))
        w = torch.zeros((Cin, Cout, 2, 2), requires_grad=True)
        F.conv_transpose2d(x, w, stride=2).backward(dout)
        eng = Engine(None)
        out = torch.empty((Cin, Cout, 2, 2), device=dev)
        # dout lives in the second half of a concat buffer, as in the decoder
        eng.wgrad(N, H, W, to_view(x, dev), to_view(dout, dev, ld=2 * Cout, off=Cout, fill=9.0), 4, 0, True, out)
        torch.cuda.synchronize()
        err = rel_l2(out, w.grad)
        assert err < 1e-4, f"convT wgrad {Cin}->{Cout}: rel-L2 {err}"


def test_wgrad_is_deterministic():
    from rbunet.engine import Engine
    dev = torch.device("cuda:0")
    x = to_view(_rand((4, 64, 32, 32), 5), dev)
    dy = to_view(_rand((4, 64, 32, 32), 6), dev)
    eng = Engine(None)
    a = torch.empty((64, 64, 3, 3), device=dev)
    b = torch.empty((64, 64, 3, 3), device=dev)
    eng.wgrad(4, 32, 32, dy, x, 9, 1, False, a)
    eng.wgrad(4, 32, 32, dy, x, 9, 1, False, b)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
