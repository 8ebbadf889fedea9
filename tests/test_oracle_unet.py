"""CPU: the plain-UNet oracle (oracle/unet_ref.py) against the golden vectors produced by the unmodified reference
class train_water_segmentation.UNet (oracle/make_golden.py: unet_golden)."""
import os

import numpy as np
import torch

from oracle import robust_unet_ref as R
from oracle import unet_ref as U


def _setup(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_c3_32x32.npz"))
    sd = R.synthetic_state_dict(U.unet_shapes(3, 2), seed=11)
    for k in sd:
        if k.endswith(".weight") and sd[k].dim() == 1:
            sd[k] = sd[k].abs()
    x, y = R.synthetic_inputs(2, 3, 32, 32, seed=321, blobby=True)
    return g, sd, x, y[:, 0].long()


def test_unet_eval_matches_reference_golden(golden_dir):
    g, sd, x, t = _setup(golden_dir)
    with torch.no_grad():
        z = U.unet_forward(sd, x, training=False)
    np.testing.assert_allclose(z.numpy(), g["logits_eval"], rtol=0, atol=2e-4)
    assert abs(U.ce_loss(z, t).item() - float(g["loss_eval"])) < 1e-4
    counts = U.argmax_counts(z.numpy(), t.numpy())
    acc, iou = U.batch_accuracy_iou(counts)
    assert abs(acc - float(g["accuracy_eval"])) < 1e-6 and abs(iou - float(g["iou_eval"])) < 1e-6


def test_unet_train_step_matches_reference_golden(golden_dir):
    g, sd, x, t = _setup(golden_dir)
    names = [str(n) for n in g["param_names"]]
    for n in names:
        sd[n].requires_grad_(True)
    z = U.unet_forward(sd, x, training=True, new_buffers={})
    loss = U.ce_loss(z, t)
    loss.backward()
    np.testing.assert_allclose(z.detach().numpy(), g["logits_train"], rtol=0, atol=2e-4)
    assert abs(loss.item() - float(g["loss_train"])) < 1e-5
    for i, n in enumerate(names):
        s = g["grad_summary"][i]
        gr = sd[n].grad.double().flatten()
        tol = 2e-4 * max(s[0], 1e-6) + 1e-7
        assert abs(gr.norm().item() - s[0]) < tol, n
        np.testing.assert_allclose(gr[:8].numpy(), s[2:2 + min(8, gr.numel())], rtol=0, atol=tol, err_msg=n)
