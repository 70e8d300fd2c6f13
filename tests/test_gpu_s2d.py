"""Space-to-depth input of the stem convolution (csrc/s2d_stem.cu, fusion.space_to_depth_stem)."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
CL = torch.channels_last


def torch_s2d(x, p):
    n, c, h, w = x.shape
    xp = F.pad(x, (p, p, p, p))
    hs, ws = (h + 2 * p) // 2, (w + 2 * p) // 2
    return xp.reshape(n, c, hs, 2, ws, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, hs, ws)


@pytest.mark.parametrize("shape,pad", [((4, 3, 224, 224), 3), ((3, 3, 32, 28), 3), ((2, 1, 6, 8), 1), ((1, 4, 2, 2), 0), ((2, 2, 6, 8), 1), ((3, 5, 4, 6), 2),
                                       ((5, 3, 10, 12), 2), ((2, 3, 2, 4), 3)])
def test_relayout_is_the_exact_gather_both_ways(shape, pad):
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    xs = ops.s2d_stem_forward(x, pad)
    ref = torch_s2d(x, pad)
    assert xs.shape == ref.shape and xs.is_contiguous(memory_format=CL) and torch.equal(xs, ref)
    gxs = torch.randn(ref.shape, generator=g).to(DEV).contiguous(memory_format=CL)
    xr = x.clone().requires_grad_(True)
    torch_s2d(xr, pad).backward(gxs)
    gx = ops.s2d_stem_backward(gxs, x.shape, pad)
    assert gx.is_contiguous(memory_format=CL) and torch.equal(gx, xr.grad)


@pytest.mark.parametrize("shape,pad,cpad", [((4, 3, 224, 224), 3, 16), ((3, 3, 32, 28), 3, 16), ((2, 1, 6, 8), 1, 8), ((2, 2, 6, 8), 1, 12),
                                            ((3, 5, 4, 6), 2, 32)])
def test_relayout_with_zero_channels_behind_the_real_ones(shape, pad, cpad):
    """``cpad``: the space-to-depth image padded to a channel count cuDNN takes without converting (12 -> 16): the real
    channels are the exact gather, the padding is zero, the backward ignores whatever gradient the padding carries."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + cpad)
    x = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    c4 = 4 * shape[1]
    xs = ops.s2d_stem_forward(x, pad, cpad=cpad)
    ref = torch_s2d(x, pad)
    assert xs.shape[1] == cpad and xs.is_contiguous(memory_format=CL)
    assert torch.equal(xs[:, :c4], ref) and not xs[:, c4:].any()
    gxs = torch.randn(xs.shape, generator=g).to(DEV).contiguous(memory_format=CL)
    xr = x.clone().requires_grad_(True)
    torch_s2d(xr, pad).backward(gxs[:, :c4])
    assert torch.equal(ops.s2d_stem_backward(gxs, x.shape, pad), xr.grad)
    with pytest.raises(RuntimeError, match="cpad"):
        ops.s2d_stem_forward(x, pad, cpad=c4 - 4 if c4 > 4 else 2)


def test_relayout_rejects_odd_extents_and_nchw():
    from ood_dfq_b200 import ops
    assert not ops.s2d_stem_supported(torch.zeros(1, 3, 7, 8, device=DEV).contiguous(memory_format=CL), 3)
    assert not ops.s2d_stem_supported(torch.zeros(2, 3, 8, 8, device=DEV), 3)
    with pytest.raises(RuntimeError):
        ops.s2d_stem_forward(torch.zeros(2, 3, 8, 8, device=DEV), 3)


@pytest.mark.parametrize("quant", [False, True])
def test_stem_convolution_is_unchanged(quant):
    """conv(x, w, stride 2, padding 3) == conv(s2d(x), re-indexed w, stride 1): values and all gradients."""
    from ood_dfq_b200 import fusion
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
    if quant:
        q = qm.Quant_Conv2d(4)
        q.set_param(conv)
        conv = q
    plain = torch.nn.Sequential(conv).to(DEV).to(memory_format=CL)
    s2d = copy.deepcopy(plain)
    x = torch.randn(4, 3, 64, 64).to(DEV).contiguous(memory_format=CL)
    assert fusion.space_to_depth_stem(s2d, x) == 1 and fusion.space_to_depth_stem(s2d, x) == 0
    assert isinstance(s2d[0], type(plain[0])) and list(s2d.state_dict()) == list(plain.state_dict())
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = plain(a), s2d(b)
    assert ya.shape == yb.shape and torch.allclose(ya, yb, rtol=1e-4, atol=1e-5)
    go = torch.randn_like(ya)
    ya.backward(go)
    yb.backward(go)
    assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(plain[0].weight.grad, s2d[0].weight.grad, rtol=1e-3, atol=1e-3)
    # an NCHW batch takes the class's own forward
    with torch.no_grad():
        assert torch.allclose(s2d(x.contiguous()), ya.detach(), rtol=1e-4, atol=1e-5)


def test_cache_follows_the_data():
    from ood_dfq_b200 import fusion
    fusion._S2DCache.clear()
    x = torch.randn(2, 3, 8, 8, device=DEV).contiguous(memory_format=CL)
    a = fusion._S2DCache.get(x, 3)
    assert fusion._S2DCache.get(x.detach(), 3) is a                   # an alias of the same data: reused
    x.add_(1.0)                                                        # in-place write: recomputed
    b = fusion._S2DCache.get(x, 3)
    assert b is not a and torch.equal(b, torch_s2d(x, 3))
    xr = x.clone().requires_grad_(True)
    c = fusion._S2DCache.get(xr, 3)
    assert c.requires_grad
    with torch.no_grad():
        d = fusion._S2DCache.get(xr, 3)
    assert d is not c and not d.requires_grad                          # no graph wanted: not the tracked tensor
    fusion._S2DCache.clear()


def test_qat_step_with_s2d_stem_matches_step_without():
    from ood_dfq_b200 import fusion, nets, step, surgery
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True

    def build(s2d):
        torch.manual_seed(1)
        teacher = nets.resnet18_imagenet(num_classes=10)
        nets.perturb_bn_stats(teacher)
        student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4).to(DEV).to(memory_format=CL)
        teacher = teacher.to(DEV).to(memory_format=CL)
        g = torch.Generator().manual_seed(2)
        xs = [torch.randn(4, 3, 224, 224, generator=g).to(DEV).contiguous(memory_format=CL) for _ in range(3)]
        with torch.no_grad():
            for x in xs[:2]:
                student(x)
        surgery.freeze_model(student)
        for m in (student, teacher):
            fusion.fuse_eval_bn(m, xs[0][:2])
            fusion.fuse_residual_tails(m, xs[0][:2])
            if s2d:
                assert fusion.space_to_depth_stem(m, xs[0][:2]) == 1
        return student, step.QATStep(student, teacher, lr=1e-5, unit_types=(nets.ResUnit,)), xs

    s0, q0, xs = build(False)
    s1, q1, _ = build(True)
    for x in xs:
        l0, l1 = q0(x), q1(x)
        assert l0.item() == l0.item() and abs(l0.item() - l1.item()) <= 2e-3 * abs(l0.item()) + 1e-5
    for (n0, p0), (n1, p1) in zip(s0.named_parameters(), s1.named_parameters()):
        assert torch.allclose(p0, p1, rtol=1e-3, atol=1e-5), n0
    fusion._S2DCache.clear()
    torch.backends.cudnn.deterministic = False
