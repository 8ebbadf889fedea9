"""Helpers shared by the GPU parity tests."""
import torch


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def to_view(x_nchw, dev, ld=None, off=0, fill=0.0):
    """NCHW fp32 (CPU) -> rbunet View on an NHWC bf16 device buffer (optionally a channel slice of a wider buffer)."""
    from rbunet import View
    n, c, h, w = x_nchw.shape
    ld = ld or c
    buf = torch.full((n, h, w, ld), fill, dtype=torch.bfloat16)
    buf[..., off:off + c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return View(buf.to(dev), off, c)


def from_view(v):
    """View -> NCHW fp32 CPU tensor."""
    return v.dense().float().cpu().permute(0, 3, 1, 2).contiguous()


def bf16r(t):
    return t.to(torch.bfloat16).float()


class Report:
    """Collects (name, error, tolerance) rows and fails once at the end with the full table (a GPU round trip
    is expensive: one run should show every deviation, not only the first)."""

    def __init__(self):
        self.rows = []

    def check(self, name, got, ref, tol):
        e = rel_l2(got.reshape(ref.shape) if hasattr(got, "reshape") else got, ref)
        self.rows.append((name, e, tol))
        return e

    def finish(self):
        bad = [r for r in self.rows if not (r[1] < r[2])]
        table = "\n".join(f"{'FAIL' if not (e < t) else 'ok  '} {n:48s} {e:10.3e} (tol {t:.1e})" for n, e, t in self.rows)
        print(table)
        assert not bad, "deviations above tolerance:\n" + table
