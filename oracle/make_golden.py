"""TEST INFRASTRUCTURE — generate `tests/golden/*.npz` by running the UNMODIFIED reference
(`/root/reference/Main_Final.py`) in the build container on deterministic synthetic weights and
inputs (integer-hash generated, see `oracle/robust_unet_ref.py`).  Run:

    python oracle/make_golden.py

The fixtures pin the oracle restatement (and through it the CUDA path) to the reference's own
outputs; the reference cannot travel to the GPU box, the fixtures do.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import robust_unet_ref as R            # noqa: E402
from oracle.load_reference import load_reference   # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _MaskQueue:
    """Replaces nn.Dropout2d.forward so the reference consumes the given channel masks in
    execution order (inc, down1, down2, down3, bottleneck.2, dec4..dec1)."""

    def __init__(self):
        self.queue = []

    def install(self):
        q = self

        def fwd(self_mod, x):
            if not self_mod.training:
                return x
            return x * q.queue.pop(0).to(x.dtype)

        self._orig = nn.Dropout2d.forward
        nn.Dropout2d.forward = fwd

    def remove(self):
        nn.Dropout2d.forward = self._orig


def grad_summary(t: torch.Tensor) -> np.ndarray:
    f = t.detach().double().flatten()
    head = f[:8].numpy()
    if head.size < 8:
        head = np.pad(head, (0, 8 - head.size))
    return np.concatenate([[f.norm().item(), f.sum().item()], head])


def model_golden(MF, tag, n_channels, base, batch, h, w, w_dice=0.0):
    shapes = R.robust_unet_shapes(n_channels, 1, base)
    torch.manual_seed(0)
    model = MF.RobustUNet(n_channels, 1, base)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), "state_dict schema drifted"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    sd = R.synthetic_state_dict(shapes, seed=0)
    model.load_state_dict(sd)
    x, y = R.synthetic_inputs(batch, n_channels, h, w, seed=123, blobby=True)
    masks = R.synthetic_drop_masks(batch, base, seed=7)

    mq = _MaskQueue()
    mq.install()
    try:
        model.train()
        mq.queue = [masks[n] for n in R.RESBLOCKS]
        p = model(x)
        loss = R.bce_dice_loss(p, y, 1.0, w_dice) if w_dice else nn.BCELoss()(p, y)
        loss.backward()
    finally:
        mq.remove()
    out = {"probs_train": p.detach().numpy(), "loss_train": np.float64(loss.item())}
    names = [k for k, _ in model.named_parameters()]
    out["param_names"] = np.array(names)
    out["grad_summary"] = np.stack([grad_summary(prm.grad) for _, prm in model.named_parameters()])
    new_sd = model.state_dict()
    bnames = [k for k in new_sd if k.endswith("running_mean") or k.endswith("running_var")]
    out["buffer_names"] = np.array(bnames)
    out["buffer_summary"] = np.stack([grad_summary(new_sd[k]) for k in bnames])
    # eval forward with the *original* synthetic weights/buffers
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        pe = model(x)
    out["probs_eval"] = pe.numpy()
    out["loss_eval"] = np.float64(nn.BCELoss()(pe, y).item())
    ev = MF.ModelEvaluator(torch.device("cpu"))
    mets = [ev.calculate_metrics(pe[i, 0], y[i, 0]) for i in range(batch)]
    keys = ["accuracy", "iou", "precision", "recall", "f1_score"]
    out["metric_keys"] = np.array(keys)
    out["metrics_eval"] = np.array([[float(m[k]) for k in keys] for m in mets], dtype=np.float64)
    out["config"] = np.array([n_channels, base, batch, h, w])
    out["w_dice"] = np.float64(w_dice)
    np.savez_compressed(os.path.join(OUT, f"model_{tag}.npz"), **out)
    print(tag, "loss", loss.item(), "eval loss", out["loss_eval"], "mean p", p.mean().item())


def module_goldens(MF):
    """Full-tensor goldens for single modules (small shapes)."""
    out = {}
    torch.manual_seed(0)
    # ResidualBlock 16 -> 32 with projection shortcut, train mode, dropout mask injected
    shapes = {}
    full = R.robust_unet_shapes(3, 1, 16)
    for k, v in full.items():
        if k.startswith("down1.1."):
            shapes[k[len("down1.1."):]] = v
    sd = R.synthetic_state_dict(shapes, seed=3)
    blk = MF.ResidualBlock(16, 32, dropout_rate=0.1)
    blk.load_state_dict(sd)
    x, _ = R.synthetic_inputs(2, 16, 8, 8, seed=11)
    x.requires_grad_(True)
    u = R._hash_uniform(2 * 32, 5)
    mask = torch.from_numpy(((u >= 0.1).astype(np.float32) / np.float32(0.9)).reshape(2, 32, 1, 1))
    mq = _MaskQueue()
    mq.install()
    try:
        blk.train()
        mq.queue = [mask]
        o = blk(x)
        go, _ = R.synthetic_inputs(2, 32, 8, 8, seed=12)
        o.backward(go)
    finally:
        mq.remove()
    out["rb_x"] = x.detach().numpy()
    out["rb_mask"] = mask.numpy()
    out["rb_out"] = o.detach().numpy()
    out["rb_gout"] = go.numpy()
    out["rb_gx"] = x.grad.numpy()
    for k, prm in blk.named_parameters():
        out["rb_g_" + k] = prm.grad.numpy()
    # AttentionGate(32, 32, 16)
    shapes = {k[len("att1."):]: v for k, v in R.robust_unet_shapes(3, 1, 32).items() if k.startswith("att1.")}
    sd = R.synthetic_state_dict(shapes, seed=4)
    gate = MF.AttentionGate(32, 32, 16)
    gate.load_state_dict(sd)
    gate.train()
    g, _ = R.synthetic_inputs(2, 32, 8, 8, seed=21)
    xs, _ = R.synthetic_inputs(2, 32, 8, 8, seed=22)
    g.requires_grad_(True)
    xs.requires_grad_(True)
    o = gate(g, xs)
    go, _ = R.synthetic_inputs(2, 32, 8, 8, seed=23)
    o.backward(go)
    out.update(ag_g=g.detach().numpy(), ag_x=xs.detach().numpy(), ag_out=o.detach().numpy(),
               ag_gout=go.numpy(), ag_gg=g.grad.numpy(), ag_gx=xs.grad.numpy())
    for k, prm in gate.named_parameters():
        out["ag_g_" + k] = prm.grad.numpy()
    # DilatedBlock(32 -> 64)
    shapes = {k[len("bottleneck.1."):]: v for k, v in R.robust_unet_shapes(3, 1, 4).items()
              if k.startswith("bottleneck.1.")}
    sd = R.synthetic_state_dict(shapes, seed=5)
    db = MF.DilatedBlock(32, 64)
    db.load_state_dict(sd)
    db.train()
    xd, _ = R.synthetic_inputs(2, 32, 8, 8, seed=31)
    xd.requires_grad_(True)
    o = db(xd)
    go, _ = R.synthetic_inputs(2, 64, 8, 8, seed=32)
    o.backward(go)
    out.update(db_x=xd.detach().numpy(), db_out=o.detach().numpy(), db_gout=go.numpy(),
               db_gx=xd.grad.numpy())
    for k, prm in db.named_parameters():
        out["db_g_" + k] = prm.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "modules.npz"), **out)


def metric_goldens(MF):
    """calculate_metrics edge cases (Main_Final.py:519-547): pred == 0.5 exactly, all-empty,
    all-full, random."""
    ev = MF.ModelEvaluator(torch.device("cpu"))
    cases = {}
    rng = np.random.RandomState(0)
    p = rng.rand(4, 16, 16).astype(np.float32)
    p[0, :4] = 0.5                                 # strict threshold: 0.5 is negative
    t = (rng.rand(4, 16, 16) > 0.5).astype(np.float32)
    p[2] = 0.0
    t[2] = 0.0                                     # empty / empty
    p[3] = 1.0
    t[3] = 1.0                                     # full / full
    keys = ["accuracy", "iou", "precision", "recall", "f1_score"]
    m = [ev.calculate_metrics(torch.from_numpy(p[i]), torch.from_numpy(t[i])) for i in range(4)]
    cases["pred"] = p
    cases["target"] = t
    cases["metric_keys"] = np.array(keys)
    cases["metrics"] = np.array([[float(mm[k]) for k in keys] for mm in m], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **cases)


def unet_golden():
    """Plain 2-class U-Net (train_water_segmentation.UNet): eval logits, one CrossEntropy train step."""
    import train_water_segmentation as T     # importable once load_reference() has installed the stubs
    from oracle import unet_ref as U
    shapes = U.unet_shapes(3, 2)
    torch.manual_seed(0)
    model = T.UNet(3, 2)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(shapes.keys()), "UNet state_dict schema drifted"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    sd = R.synthetic_state_dict(shapes, seed=11)
    # positive BatchNorm scales keep this reference close to a trained network (defaults are gamma = 1)
    for k in sd:
        if k.endswith(".weight") and sd[k].dim() == 1:
            sd[k] = sd[k].abs()
    model.load_state_dict(sd)
    x, y = R.synthetic_inputs(2, 3, 32, 32, seed=321, blobby=True)
    t = y[:, 0].long()
    model.train()
    z = model(x)
    loss = nn.CrossEntropyLoss()(z, t)
    loss.backward()
    out = {"logits_train": z.detach().numpy(), "loss_train": np.float64(loss.item())}
    out["param_names"] = np.array([k for k, _ in model.named_parameters()])
    out["grad_summary"] = np.stack([grad_summary(p.grad) for _, p in model.named_parameters()])
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        ze = model(x)
    out["logits_eval"] = ze.numpy()
    out["loss_eval"] = np.float64(nn.CrossEntropyLoss()(ze, t).item())
    pred = torch.argmax(ze, dim=1)
    out["accuracy_eval"] = np.float64((pred == t).float().mean().item())
    tr = T.WaterSegmentationTrainer.calculate_iou(None, pred == 1, t == 1)
    out["iou_eval"] = np.float64(float(tr))
    np.savez_compressed(os.path.join(OUT, "unet_c3_32x32.npz"), **out)
    print("unet loss", loss.item(), "eval loss", out["loss_eval"], "acc", out["accuracy_eval"], "iou", out["iou_eval"])


def _scene(rng, h, w, c, hi, dtype):
    """Synthetic satellite-like bands: smooth low-frequency field + speckle, skewed like digital numbers."""
    low = rng.random((h // 8 + 2, w // 8 + 2, c))
    up = np.kron(low, np.ones((8, 8, 1)))[:h, :w, :]
    v = (up ** 2) * hi * 0.6 + rng.gamma(2.0, hi / 40.0, size=(h, w, c))
    return np.clip(v, 0, hi - 1).astype(dtype)


def imageops_golden():
    """the reference's own enhance_image (tif_to_image.py:139-171, called unbound -- it
    does not use self) and the OpenCV calls of predict_coastline.py:595-602 on small seeded inputs."""
    import importlib
    import cv2
    T = importlib.import_module("tif_to_image")      # osgeo is stubbed by load_reference()
    conv = T.TIFToImageConverter.__new__(T.TIFToImageConverter)
    rng = np.random.default_rng(2026)
    out = {"cv2_version": np.array(cv2.__version__)}
    cases = [("u16", np.uint16, 30000, (48, 40, 3)), ("u8", np.uint8, 256, (33, 47, 3)), ("u16full", np.uint16, 65536, (24, 56, 4)),
             ("tiny", np.uint8, 256, (3, 5, 3))]
    for tag, dt, hi, shape in cases:
        rgb = _scene(rng, *shape, hi, dt)
        out[f"enh_{tag}_in"] = rgb
        for ew in (0, 1):
            out[f"enh_{tag}_out{ew}"] = conv.enhance_image(rgb, bool(ew)).astype(np.uint8)
        out[f"enh_{tag}_pct"] = np.array([[np.percentile(rgb[:, :, i], q) for q in (2, 98)] for i in range(shape[2])])
    for k in (1, 2, 3, 5, 8, 20, 31):
        out[f"ellipse_{k}"] = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    for tag, shape in (("a", (40, 56)), ("b", (16, 16)), ("c", (5, 70))):
        blob = _scene(rng, shape[0], shape[1], 1, 256, np.uint8)[:, :, 0]
        mask = ((blob > np.median(blob)) * 255).astype(np.uint8)
        mask01 = (mask // 255).astype(np.uint8)
        out[f"mask_{tag}"] = mask
        for k in (5, 20, 3):
            kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
            d = cv2.dilate(mask, kern, iterations=1)
            out[f"coast_{tag}_k{k}"] = d - mask
            out[f"coast01_{tag}_k{k}"] = cv2.dilate(mask01, kern, iterations=1) - mask01
        out[f"gray_{tag}"] = blob
        out[f"graydil_{tag}_k5"] = cv2.dilate(blob, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), iterations=1)
    np.savez_compressed(os.path.join(OUT, "imageops.npz"), **out)
    print("imageops goldens:", len(out), "arrays")


def base64_goldens(MF):
    """The reference's own width (base 64, up to 1024 channels, Main_Final.py:229) at sizes whose fixtures stay small:
    probabilities in full, gradients as per-tensor (norm, sum, first 8 values) summaries of all 173 parameters."""
    model_golden(MF, "c3_b64_64x64", 3, 64, 2, 64, 64)
    model_golden(MF, "c4_b64_64x96", 4, 64, 2, 64, 96, w_dice=0.5)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    MF = load_reference()
    if sys.argv[1:] == ["base64"]:          # only the base-64 fixtures (the others are unchanged)
        base64_goldens(MF)
        return
    base64_goldens(MF)
    model_golden(MF, "c3_b16_32x32", 3, 16, 2, 32, 32)
    model_golden(MF, "c4_b16_32x48", 4, 16, 2, 32, 48, w_dice=0.5)
    module_goldens(MF)
    metric_goldens(MF)
    unet_golden()
    imageops_golden()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
