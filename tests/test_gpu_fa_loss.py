"""Feature-alignment loss of the QAT step as one kernel each way (csrc/fa_loss.cu) against the reference's own
expression: ``lam * sum_l ((F.normalize(Es_l) - F.normalize(Et_l)) ** 2).mean()`` (trainer_direct.py:325-330 over the
maps of :382-383), forward value and the gradient w.r.t. every energy, within 1e-5 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference(es, et, lam):
    fa = torch.zeros(1, dtype=torch.float64)
    for s, t in zip(es, et):
        fa = fa + (F.normalize(s) - F.normalize(t)).pow(2).mean()
    return lam * fa


def energies(shapes, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(s, generator=g).square() * scale) for s in shapes]


@pytest.mark.parametrize("shapes", [[(256, 64), (256, 64), (256, 128), (256, 128), (256, 256), (256, 256), (256, 512), (256, 512)],
                                    [(256, 16)] * 3 + [(256, 32)] * 3 + [(256, 64)] * 3, [(3, 5)], [(7, 1000), (7, 12)]])
def test_fused_fa_loss_matches_the_torch_expression(shapes):
    from ood_dfq_b200 import step
    lam = 1000.0
    es, et = energies(shapes, 1), energies(shapes, 2, 1.7)
    es64 = [e.double().requires_grad_(True) for e in es]
    et64 = [e.double().requires_grad_(True) for e in et]
    ref = reference(es64, et64, lam)
    (ref * 0.37).backward()
    es_g = [e.to(DEV).requires_grad_(True) for e in es]
    et_g = [e.to(DEV).requires_grad_(True) for e in et]
    got = step.feature_alignment_loss(es_g, et_g, lam, DEV, raw=True)
    assert got.shape == (1,)
    np.testing.assert_allclose(got.item(), ref.item(), rtol=1e-5)
    (got * 0.37).backward()
    for a, b in zip(es_g + et_g, es64 + et64):
        scale = b.grad.abs().max().item()
        np.testing.assert_allclose(a.grad.cpu().double().numpy(), b.grad.numpy(), rtol=1e-5, atol=1e-6 * scale)


def test_one_launch_each_way_and_one_sided_gradients():
    from ood_dfq_b200 import _native, step
    shapes = [(32, 64)] * 4
    es = [e.to(DEV).requires_grad_(True) for e in energies(shapes, 3)]
    et = [e.to(DEV) for e in energies(shapes, 4)]                      # teacher under no_grad: no gradient wanted
    step.feature_alignment_loss(es, et, 10.0, DEV, raw=True).backward()  # workspace allocation out of the way
    torch.cuda.synchronize()
    _native.reset_launch_count()
    loss = step.feature_alignment_loss(es, et, 10.0, DEV, raw=True)
    assert _native.launch_count() == 1
    loss.backward()
    assert _native.launch_count() == 2
    es64 = [e.detach().cpu().double().requires_grad_(True) for e in es]
    reference(es64, [e.cpu().double() for e in et], 10.0).backward()
    for a, b in zip(es, es64):
        np.testing.assert_allclose(a.grad.cpu().double().numpy() / 2, b.grad.numpy(), rtol=1e-5, atol=1e-9)   # two backwards above


def test_zero_rows_take_the_clamped_branch():
    """A row of zeros has norm 0 < eps: F.normalize divides by eps and passes no gradient through the norm."""
    from ood_dfq_b200 import step
    es = energies([(4, 8)], 5)
    es[0][1] = 0.0
    et = energies([(4, 8)], 6)
    es64, et64 = [es[0].double().requires_grad_(True)], [et[0].double().requires_grad_(True)]
    ref = reference(es64, et64, 3.0)
    ref.backward()
    a, b = es[0].to(DEV).requires_grad_(True), et[0].to(DEV).requires_grad_(True)
    got = step.feature_alignment_loss([a], [b], 3.0, DEV, raw=True)
    got.backward()
    np.testing.assert_allclose(got.item(), ref.item(), rtol=1e-5)
    # the clamped row's gradient is dA / eps = O(1e12): compare relative to its own magnitude
    for x, y in ((a, es64[0]), (b, et64[0])):
        np.testing.assert_allclose(x.grad.cpu().double().numpy(), y.grad.numpy(), rtol=1e-5, atol=1e-9)
