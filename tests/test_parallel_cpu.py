"""CPU (gloo, world_size 2): the bucketed gradient all-reduce of rbunet.parallel.GradBucketer -- bucket layout in
reverse execution order, stage-by-stage readiness, averaging -- against a direct average of the per-rank gradients."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import robust_unet_ref as R
    from rbunet.parallel import STAGE_ORDER, GradBucketer
    shapes = {k: v for k, v in R.robust_unet_shapes(3, 1, 16).items()
              if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))}
    named = [(k, torch.Size(v)) for k, v in shapes.items()]
    b = GradBucketer(named, "cpu", bucket_bytes=256 << 10)
    assert len(b.flat) > 3
    # buckets follow the order in which the engine finishes stages
    first = b.bucket_members[0][0].split(".")[0]
    last = b.bucket_members[-1][-1].split(".")[0]
    assert first == "outc" and last == "inc"
    for step in range(2):                      # two steps: the bucketer resets itself
        gen = torch.Generator().manual_seed(100 * step + rank)
        grads = {k: torch.randn(s, generator=gen) for k, s in named}
        for stage in STAGE_ORDER:              # stages become ready one at a time, like Engine.backward
            b.ready([k for k in grads if k.split(".")[0] == stage], grads)
        red = b.finish()
        want = {}
        for k, s in named:
            acc = torch.zeros(s)
            for r in range(world):
                g2 = torch.Generator().manual_seed(100 * step + r)
                allg = {kk: torch.randn(ss, generator=g2) for kk, ss in named}
                acc += allg[k]
            want[k] = acc / world
        for k, _ in named:
            assert red[k].shape == want[k].shape
            assert torch.allclose(red[k], want[k], atol=1e-6), k
    torch.save(True, os.path.join(out_dir, f"ok{rank}"))
    dist.destroy_process_group()


def test_grad_bucketer_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_reverse_execution_order_covers_all_parameters():
    sys.path.insert(0, ROOT)
    from oracle import robust_unet_ref as R
    from rbunet.parallel import reverse_execution_order
    names = [k for k in R.robust_unet_shapes(3, 1, 64) if "running" not in k and "num_batches" not in k]
    order = reverse_execution_order(names)
    assert sorted(order) == sorted(names) and len(order) == 173
    assert order[0].startswith("outc") and order[-1].startswith("inc.")
