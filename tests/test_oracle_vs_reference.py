"""CPU: the oracle restatement against the LIVE unmodified reference (imported from /root/reference with matplotlib
stubs) at the reference's own width (base 64).  Skipped where the reference tree is absent (the GPU box): there the
committed goldens under tests/golden/ pin the oracle instead (tests/test_oracle_golden.py)."""
import pytest
import torch

from oracle import robust_unet_ref as R
from oracle.load_reference import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present on this machine")


def _same(a, b):
    """Bit-identical in a quiet process; <= 1e-5 rel-L2 when other tests have changed the intra-op thread partitioning."""
    return torch.equal(a, b) or ((a.double() - b.double()).norm() <= 1e-5 * b.double().norm() + 1e-12).item()


@pytest.fixture(scope="module")
def MF():
    return load_reference()


def test_reference_init_is_bit_identical_and_state_dict_loads(MF):
    """Same constructor argument order, same parameter registration order => the same torch RNG stream => identical
    initial weights (Main_Final.py:229-288); the reference state_dict loads into the product module and back."""
    import rbunet
    for nc in (3, 4):
        torch.manual_seed(0)
        ref = MF.RobustUNet(nc, 1)
        torch.manual_seed(0)
        ours = rbunet.RobustUNet(nc, 1, 64)
        rsd, osd = ref.state_dict(), ours.state_dict()
        assert list(rsd.keys()) == list(osd.keys()) and len(rsd) == 290
        for k in rsd:
            assert rsd[k].dtype == osd[k].dtype and torch.equal(rsd[k], osd[k]), k
        ours.load_state_dict(rsd)
        ref.load_state_dict(ours.state_dict())


def test_oracle_equals_reference_at_base64_train_and_eval(MF):
    """Same torch CPU kernels in the same order at base 64, 3 x 128 x 128, reference initialisation, Dropout2d masks
    injected into both: bit-identical probabilities; gradients bit-identical when the test runs alone and within 1e-5
    (rel-L2) always -- oneDNN's backward kernels split their reductions over however many threads the process has at
    that moment, which other tests of the suite change."""
    import torch.nn as nn
    torch.manual_seed(0)
    ref = MF.RobustUNet(3, 1)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    x, y = R.synthetic_inputs(1, 3, 128, 128, seed=123, blobby=True)
    masks = R.synthetic_drop_masks(1, 64, seed=7)
    queue = [masks[n] for n in R.RESBLOCKS]
    orig = nn.Dropout2d.forward
    nn.Dropout2d.forward = lambda self, t: t * queue.pop(0) if self.training else t
    try:
        ref.train()
        p_ref = ref(x)
        nn.BCELoss()(p_ref, y).backward()
    finally:
        nn.Dropout2d.forward = orig
    names = [n for n, _ in ref.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    p = R.robust_unet_forward(sd, x, training=True, drop_masks=masks, new_buffers={})
    R.bce_loss(p, y).backward()
    assert _same(p.detach(), p_ref.detach())
    for n, prm in ref.named_parameters():
        a, b = sd[n].grad.double(), prm.grad.double()
        assert (a - b).norm() <= 1e-5 * b.norm() + 1e-7, n          # floor: biases in front of a train-mode BN have pure round-off gradients
    ref.eval()
    with torch.no_grad():
        pe_ref = ref(x)
        sd2 = {k: v.detach() for k, v in ref.state_dict().items()}
        pe = R.robust_unet_forward(sd2, x, training=False)
    assert _same(pe, pe_ref)
