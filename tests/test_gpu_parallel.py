"""GPU: rbunet.DataParallel on a single-rank NCCL group -- the bucketed gradient path (stage hooks, flat buckets,
communication stream, views as .grad) must reproduce the plain model's gradients exactly.  The multi-rank averaging
itself is covered on the CPU by tests/test_parallel_cpu.py (gloo, world size 2) and by bench.py under torchrun."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

from oracle import robust_unet_ref as R

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_data_parallel_world1_matches_plain_model():
    import rbunet
    dev = torch.device("cuda:0")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        sd = R.synthetic_state_dict(R.robust_unet_shapes(3, 1, 16), seed=0)
        x, y = R.synthetic_inputs(2, 3, 32, 32, seed=3, blobby=True)
        masks = R.synthetic_drop_masks(2, 16, seed=7)
        grads = []
        for wrap in (False, True):
            model = rbunet.RobustUNet(3, 1, 16)
            model.load_state_dict(sd)
            model.to(dev).train()
            model.engine.drop_mask_fn = lambda nm, N, C: masks[nm]
            net = rbunet.DataParallel(model, bucket_bytes=64 << 10) if wrap else model
            loss = rbunet.RobustBCEDiceLoss()(net(x.to(dev)), y.to(dev))
            loss.backward()
            torch.cuda.synchronize()
            grads.append({n: p.grad.clone() for n, p in model.named_parameters()})
            if wrap:
                assert len(net.bucketer.flat) > 4
        for n in grads[0]:
            assert torch.equal(grads[0][n], grads[1][n]), n
    finally:
        dist.destroy_process_group()
