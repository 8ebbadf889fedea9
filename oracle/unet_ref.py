"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's plain 2-class U-Net path
(`/root/reference/train_water_segmentation.py:209-288` model, `:304` CrossEntropyLoss, `:384-388` argmax accuracy / IoU;
the same network is loaded by `predict_coastline.py:255-334,390-392`).  Functional (state_dict in, tensors out), pinned by
golden vectors produced from the unmodified reference class (`oracle/make_golden.py` → `tests/golden/unet_*.npz`).
Only tests / smoke / bench may import this module; the product package never does.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .robust_unet_ref import BF16, FP32, batch_norm  # noqa: F401  (same storage model and BatchNorm restatement)

BLOCKS = ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec4", "dec3", "dec2", "dec1")


def unet_shapes(n_channels=3, n_classes=2) -> Dict[str, tuple]:
    """The 136-entry state_dict schema of train_water_segmentation.UNet in registration order (:222-246)."""
    shapes: Dict[str, tuple] = {}

    def bn(p, c):
        shapes[p + ".weight"] = (c,)
        shapes[p + ".bias"] = (c,)
        shapes[p + ".running_mean"] = (c,)
        shapes[p + ".running_var"] = (c,)
        shapes[p + ".num_batches_tracked"] = ()

    def block(p, ci, co):                      # conv_block (:251-260): conv, BN, ReLU, conv, BN, ReLU
        shapes[p + ".0.weight"] = (co, ci, 3, 3)
        shapes[p + ".0.bias"] = (co,)
        bn(p + ".1", co)
        shapes[p + ".3.weight"] = (co, co, 3, 3)
        shapes[p + ".3.bias"] = (co,)
        bn(p + ".4", co)

    block("enc1", n_channels, 64)
    block("enc2", 64, 128)
    block("enc3", 128, 256)
    block("enc4", 256, 512)
    block("bottleneck", 512, 1024)
    for k, c in ((4, 512), (3, 256), (2, 128), (1, 64)):
        shapes[f"upconv{k}.weight"] = (2 * c, c, 2, 2)
        shapes[f"upconv{k}.bias"] = (c,)
        block(f"dec{k}", 2 * c, c)
    shapes["final.weight"] = (n_classes, 64, 1, 1)
    shapes["final.bias"] = (n_classes,)
    return shapes


def conv_block(sd, p, x, training, new_buffers=None, st=FP32):
    """conv_block (train_water_segmentation.py:251-260)."""
    y = st.act(F.conv2d(x, st.weight(sd[p + ".0.weight"]), sd[p + ".0.bias"], padding=1))
    a = st.act(F.relu(batch_norm(sd, p + ".1", y, training, new_buffers)))
    y = st.act(F.conv2d(a, st.weight(sd[p + ".3.weight"]), sd[p + ".3.bias"], padding=1))
    return st.act(F.relu(batch_norm(sd, p + ".4", y, training, new_buffers)))


def unet_forward(sd, x, training=False, new_buffers: Optional[dict] = None, st=FP32):
    """UNet.forward (train_water_segmentation.py:262-288): concat order is [upconv, encoder] (:274,278,282,286).
    Returns logits [B, n_classes, H, W]."""
    e1 = conv_block(sd, "enc1", st.act(x), training, new_buffers, st)
    e2 = conv_block(sd, "enc2", F.max_pool2d(e1, 2), training, new_buffers, st)
    e3 = conv_block(sd, "enc3", F.max_pool2d(e2, 2), training, new_buffers, st)
    e4 = conv_block(sd, "enc4", F.max_pool2d(e3, 2), training, new_buffers, st)
    t = conv_block(sd, "bottleneck", F.max_pool2d(e4, 2), training, new_buffers, st)
    for k, skip in ((4, e4), (3, e3), (2, e2), (1, e1)):
        u = st.act(F.conv_transpose2d(t, st.weight(sd[f"upconv{k}.weight"]), sd[f"upconv{k}.bias"], stride=2))
        t = conv_block(sd, f"dec{k}", torch.cat([u, skip], dim=1), training, new_buffers, st)
    return F.conv2d(t, sd["final.weight"], sd["final.bias"])


def ce_loss(logits, target):
    """nn.CrossEntropyLoss() with class-index targets [B,H,W] (train_water_segmentation.py:304,381)."""
    return F.cross_entropy(logits, target)


def argmax_counts(logits: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Integer TP/FP/FN/TN per image of `argmax(dim=1) == 1` against `target == 1` (:384-388; ties -> class 0)."""
    pred = np.argmax(logits, axis=1).reshape(logits.shape[0], -1) == 1
    tb = target.reshape(target.shape[0], -1) == 1
    tp = (pred & tb).sum(1)
    fp = (pred & ~tb).sum(1)
    fn = (~pred & tb).sum(1)
    tn = pred.shape[1] - tp - fp - fn
    return np.stack([tp, fp, fn, tn], axis=1).astype(np.int64)


def batch_accuracy_iou(counts: np.ndarray):
    """accuracy = mean(pred == mask) and calculate_iou (:341-358) over the whole batch, from integer counts."""
    tp, fp, fn, tn = [float(v) for v in counts.sum(0)]
    union = tp + fp + fn
    return (tp + tn) / (tp + fp + fn + tn), (1.0 if union == 0 else tp / union)
