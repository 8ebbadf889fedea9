"""Diagnostic (GPU): block-by-block outputs of rbunet.RobustUNet against the bf16-storage oracle (eval or train mode) --
localises a deviation to the first block whose output differs.  Usage: python tools/layer_diff.py B H W [train]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rbunet                                       # noqa: E402
from gpu_util import from_view, rel_l2              # noqa: E402
from oracle import robust_unet_ref as R            # noqa: E402


def main():
    B, H, W = (int(v) for v in sys.argv[1:4])
    training = len(sys.argv) > 4 and sys.argv[4] == "train"
    base = int(os.environ.get("BASE", "16"))
    dev = torch.device("cuda:0")
    sd = R.synthetic_state_dict(R.robust_unet_shapes(3, 1, base), seed=0)
    x, y = R.synthetic_inputs(B, 3, H, W, seed=31, blobby=True)
    masks = R.synthetic_drop_masks(B, base, seed=7)
    model = rbunet.RobustUNet(3, 1, base)
    model.load_state_dict(sd)
    model.to(dev).train(training)
    model.engine.drop_mask_fn = lambda nm, N, C: masks[nm]
    with torch.no_grad():
        probs, S = model.engine.forward(x.to(dev), training, True)
    torch.cuda.synchronize()
    st = R.BF16
    dm = masks if training else {}
    with torch.no_grad():
        ref = {}
        t = st.act(x)
        ref["inc"] = x1 = R.residual_block(sd, "inc", t, training, dm.get("inc"), None, st)
        ref["down1"] = x2 = R.residual_block(sd, "down1.1", F.max_pool2d(x1, 2), training, dm.get("down1.1"), None, st)
        ref["down2"] = x3 = R.residual_block(sd, "down2.1", F.max_pool2d(x2, 2), training, dm.get("down2.1"), None, st)
        ref["down3"] = x4 = R.residual_block(sd, "down3.1", F.max_pool2d(x3, 2), training, dm.get("down3.1"), None, st)
        ref["dil"] = x5a = R.dilated_block(sd, "bottleneck.1", F.max_pool2d(x4, 2), training, None, st)
        ref["bott"] = t = R.residual_block(sd, "bottleneck.2", x5a, training, dm.get("bottleneck.2"), None, st)
        for k, skip in ((4, x4), (3, x3), (2, x2), (1, x1)):
            up = st.act(F.conv_transpose2d(t, st.weight(sd[f"up{k}.weight"]), sd[f"up{k}.bias"], stride=2))
            ref[f"up{k}"] = up
            att = R.attention_gate(sd, f"att{k}", up, skip, training, None, st)
            ref[f"att{k}"] = att
            ref[f"dec{k}"] = t = R.residual_block(sd, f"dec{k}", torch.cat([att, up], 1), training, dm.get(f"dec{k}"), None, st)
        pr = torch.sigmoid(F.conv2d(t, sd["outc.0.weight"], sd["outc.0.bias"]))
    print(f"B={B} {H}x{W} base {base} {'train' if training else 'eval'}")
    for name in ("inc", "down1", "down2", "down3", "bott"):
        print(f"  {name:6s} out rel-L2 {rel_l2(from_view(S[name]['out']), ref[name]):.3e}")
    print(f"  dil    out rel-L2 {rel_l2(from_view(S['bott']['x']), ref['dil']):.3e}")
    for k in (4, 3, 2, 1):
        sdk = S[f"dec{k}"]
        cat = sdk["x"]
        C = cat.C // 2
        print(f"  att{k}   out rel-L2 {rel_l2(from_view(cat.slice(0, C)), ref[f'att{k}']):.3e}   up{k} {rel_l2(from_view(cat.slice(C, C)), ref[f'up{k}']):.3e}"
              f"   dec{k} {rel_l2(from_view(sdk['out']), ref[f'dec{k}']):.3e}")
    print(f"  probs rel-L2 {rel_l2(probs, pr):.3e}")


if __name__ == "__main__":
    main()
