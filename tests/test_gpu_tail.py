"""Residual-unit tail (csrc/res_tail.cu): BN1(x1) + identity -> ReLU -> QuantAct (+ feature-alignment energy) as
one kernel each way, against the chain of kernels / ATen ops it replaces and against the CPU oracle modules."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
CL = torch.channels_last
SHAPES = [(8, 64, 56, 56), (16, 128, 28, 28), (16, 256, 14, 14), (32, 512, 7, 7), (3, 8, 5, 3), (2, 12, 9, 1),
          (1, 4, 1, 1), (2, 1024, 2, 2), (300, 16, 8, 8)]


def make_bn(c, g):
    return ((torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.3).to(DEV),
            (torch.randn(c, generator=g) * 0.2).to(DEV), (torch.rand(c, generator=g) + 0.4).to(DEV), 1e-5)


def chain_forward(ops, x1, r, bn1, bn2, fq):
    """The unfused sequence: fused-BN kernel(s), ATen add, ReLU+QuantAct kernel, channel-energy kernel."""
    z1 = ops.bn_eval_forward(x1, *bn1)
    rid = ops.bn_eval_forward(r, *bn2) if bn2 is not None else r
    s = z1 + rid
    y = ops.fake_quant(s, fq[0], fq[1], fq[2], relu_first=True) if fq is not None else torch.relu(s)
    return z1, s, y, ops.channel_energy_forward(z1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("idbn", [False, True])
@pytest.mark.parametrize("k", [4, 2, 0])
def test_tail_matches_the_kernel_chain(shape, idbn, k):
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + 7 * k + idbn)
    x1 = (torch.randn(shape, generator=g) * 1.3).to(DEV).contiguous(memory_format=CL)
    r = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    if not idbn:
        r = torch.relu(r)                  # a plain identity is the previous unit's post-ReLU output
    bn1, bn2 = make_bn(shape[1], g), (make_bn(shape[1], g) if idbn else None)
    fq = (k, torch.zeros(1, device=DEV), torch.full((1,), 2.3, device=DEV)) if k else None
    z1, s, y_ref, e_ref = chain_forward(ops, x1, r, bn1, bn2, fq)
    y, e = ops.res_tail_forward(x1, r, bn1, bn2, fq=fq, want_energy=True)
    assert y.is_contiguous(memory_format=CL) and torch.equal(y, y_ref)
    np.testing.assert_allclose(e.cpu().numpy(), e_ref.cpu().numpy(), rtol=2e-5, atol=1e-7)
    y2, e2 = ops.res_tail_forward(x1, r, bn1, bn2, fq=fq, want_energy=False)
    assert e2 is None and torch.equal(y2, y)

    gy = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    ge = torch.randn(shape[:2], generator=g).to(DEV)
    for with_energy in (True, False):
        gm = torch.ops.aten.threshold_backward(gy, s, 0)
        dz = gm + ops.channel_energy_backward(z1, ge) if with_energy else gm
        gx1_ref, dw1_ref, db1_ref = ops.bn_eval_backward(x1, dz, *bn1, relu=False)
        if idbn:
            gr_ref, dw2_ref, db2_ref = ops.bn_eval_backward(r, gm, *bn2, relu=False)
        else:
            gr_ref = gm
        gx1, gr, dw1, db1, dw2, db2 = ops.res_tail_backward(gy, ge if with_energy else None, x1, r, bn1, bn2)
        assert torch.equal(gx1, gx1_ref) and torch.equal(gr, gr_ref)
        n_red = x1.numel() / shape[1]
        tol = dict(rtol=2e-4, atol=3e-5 * n_red ** 0.5)
        np.testing.assert_allclose(dw1.cpu().numpy(), dw1_ref.cpu().numpy(), **tol)
        np.testing.assert_allclose(db1.cpu().numpy(), db1_ref.cpu().numpy(), **tol)
        if idbn:
            np.testing.assert_allclose(dw2.cpu().numpy(), dw2_ref.cpu().numpy(), **tol)
            np.testing.assert_allclose(db2.cpu().numpy(), db2_ref.cpu().numpy(), **tol)
        else:
            assert dw2 is None and db2 is None
        gx1b, grb, dw1b, _, _, _ = ops.res_tail_backward(gy, ge if with_energy else None, x1, r, bn1, bn2,
                                                         want_param_grads=False)
        assert dw1b is None and torch.equal(gx1b, gx1) and torch.equal(grb, gr)
        # with the forward's ReLU mask the backward no longer re-derives it: same bits, and the tensors the variant
        # does not need may be withheld altogether
        y3, _, mask = ops.res_tail_forward(x1, r, bn1, bn2, fq=fq, want_energy=False, want_mask=True)
        assert torch.equal(y3, y) and mask.dtype == torch.uint8 and mask.numel() == x1.numel() // 4
        m = ops.res_tail_backward(gy, ge if with_energy else None, x1, r if idbn else None, bn1, bn2, mask=mask)
        assert torch.equal(m[0], gx1) and torch.equal(m[1], gr)
        assert torch.equal(m[2], dw1) and torch.equal(m[3], db1)
        if idbn:
            assert torch.equal(m[4], dw2) and torch.equal(m[5], db2)
        m = ops.res_tail_backward(gy, ge if with_energy else None, x1 if with_energy else None, None, bn1, bn2,
                                  want_param_grads=False, mask=mask)
        assert torch.equal(m[0], gx1) and torch.equal(m[1], gr) and m[2] is None
    with pytest.raises(RuntimeError, match="needs x1 / r"):
        ops.res_tail_backward(gy, None, x1, None, bn1, bn2)                      # no mask: r is indispensable


@pytest.mark.parametrize("idbn", [False, True])
@pytest.mark.parametrize("k", [4, 0])
def test_tail_against_the_cpu_oracle(idbn, k):
    """Directly against the unfused torch-CPU modules.  The kernel's affine rounds once (FFMA) where ATen's
    batch_norm rounds more often, so the pre-activation may differ in the last ulp: outputs are then either equal
    or -- where that ulp crosses a rounding boundary of the quantiser -- exactly one quantisation step apart, and
    the ReLU mask of the gradient may differ only where the pre-activation is (numerically) zero."""
    from ood_dfq_b200 import ops
    from oracle import fused_torch
    g = torch.Generator().manual_seed(21 + idbn + k)
    shape = (6, 32, 14, 14)
    x1 = (torch.randn(shape, generator=g) * 1.3).requires_grad_(True)
    r = torch.randn(shape, generator=g)
    r = (r if idbn else torch.relu(r)).requires_grad_(True)
    bns = [tuple(t.cpu() if isinstance(t, torch.Tensor) else t for t in make_bn(32, g)) for _ in range(2)]
    bn1, bn2 = bns[0], (bns[1] if idbn else None)
    lo, hi = torch.zeros(1), torch.full((1,), 2.3)
    y_ref, e_ref, s_ref = fused_torch.residual_tail(x1, r, bn1, bn2, k, lo, hi)
    gy = torch.randn(shape, generator=g)
    ge = torch.randn(shape[:2], generator=g)
    (y_ref * gy).sum().backward(retain_graph=True)
    gx1_y, gr_y = x1.grad.clone(), r.grad.clone()
    x1.grad = None
    (e_ref * ge).sum().backward()
    gx1_ref = gx1_y + x1.grad

    def dev(t):
        return t.detach().to(DEV).contiguous(memory_format=CL) if t.dim() == 4 else t.to(DEV)
    d1 = tuple(dev(t) if isinstance(t, torch.Tensor) else t for t in bn1)
    d2 = None if bn2 is None else tuple(dev(t) if isinstance(t, torch.Tensor) else t for t in bn2)
    fq = (k, lo.to(DEV), hi.to(DEV)) if k else None
    y, e = ops.res_tail_forward(dev(x1), dev(r), d1, d2, fq=fq, want_energy=True)
    diff = (y.cpu() - y_ref.detach()).abs()
    step = (2.3 / (2 ** k - 1)) if k else 0.0
    close = diff <= 1e-5 * (1 + y_ref.detach().abs())
    assert (close | ((diff - step).abs() <= 1e-5)).all() and close.float().mean() > 0.999
    np.testing.assert_allclose(e.cpu().numpy(), e_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    gx1, gr, _, _, _, _ = ops.res_tail_backward(dev(gy), ge.to(DEV), dev(x1), dev(r), d1, d2)
    sure = s_ref.detach().abs() > 1e-5                       # away from the ReLU kink both masks agree
    np.testing.assert_allclose(gx1.cpu()[sure].numpy(), gx1_ref[sure].numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gr.cpu()[sure].numpy(), gr_y[sure].numpy(), rtol=1e-4, atol=1e-6)


def test_second_output_gradient_is_summed_in_the_kernel():
    """grad_y2 / grad_out2: the kernels add the two gradients exactly like autograd's accumulation would."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(11)
    shape = (6, 32, 14, 14)
    x1 = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    r = torch.relu(torch.randn(shape, generator=g)).to(DEV).contiguous(memory_format=CL)
    bn1, bn2 = make_bn(32, g), make_bn(32, g)
    ga = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    gb = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=CL)
    ge = torch.randn(shape[:2], generator=g).to(DEV)
    for b2 in (None, bn2):
        one = ops.res_tail_backward(ga + gb, ge, x1, r, bn1, b2)
        two = ops.res_tail_backward(ga, ge, x1, r, bn1, b2, grad_y2=gb)
        for a, b in zip(one, two):
            assert (a is None and b is None) or torch.equal(a, b)
    # stem
    xs = (torch.randn(4, 16, 20, 18, generator=g) * 1.4).to(DEV).contiguous(memory_format=CL)
    w, b, rm, rv, eps = make_bn(16, g)
    fq = (4, torch.zeros(1, device=DEV), torch.full((1,), 1.9, device=DEV))
    out, idx, xhat = ops.bn_pool_forward(xs, w, b, rm, rv, eps, fq=fq)
    ga = torch.randn(out.shape, generator=g).to(DEV).contiguous(memory_format=CL)
    gb = torch.randn(out.shape, generator=g).to(DEV).contiguous(memory_format=CL)
    one = ops.bn_pool_backward(ga + gb, idx, xhat, xs.shape, w, b, rm, rv, eps)
    two = ops.bn_pool_backward(ga, idx, xhat, xs.shape, w, b, rm, rv, eps, grad_out2=gb)
    for a, b_ in zip(one, two):
        assert torch.equal(a, b_)


def test_fused_units_hand_each_other_a_second_handle():
    """The producer hangs a twin of its output on the tensor; the next fused unit reads its identity through it,
    so no autograd accumulation kernel runs at the unit input -- and the gradients do not change."""
    from ood_dfq_b200 import fusion
    plain, fused, xs, _ = _pair("resnet20_cifar", CL, 32)
    units = [m for m in fused.modules() if isinstance(m, fusion._FusedUnitMixin)]
    seen = []
    hooks = [u.register_forward_hook(lambda m, i, o: seen.append((hasattr(i[0], fusion._TWIN), hasattr(o, fusion._TWIN))))
             for u in units]
    a, b = xs[1].clone().requires_grad_(True), xs[1].clone().requires_grad_(True)
    ya, yb = plain(a), fused(b)
    for h in hooks:
        h.remove()
    assert all(out for _, out in seen) and all(inp for inp, _ in seen[1:])
    ya.square().mean().backward()
    yb.square().mean().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-7 + 1e-5 * a.grad.abs().max().item())


def test_tail_rejects_what_it_cannot_run():
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 4, 4, generator=g).to(DEV)                       # NCHW
    bn = make_bn(8, g)
    assert not ops.res_tail_supported(x, x)
    with pytest.raises(RuntimeError):
        ops.res_tail_forward(x, x, bn)
    with pytest.raises(RuntimeError):
        ops.res_tail_forward(x.cpu(), x.cpu(), bn)
    xl = torch.randn(2, 6, 4, 4, generator=g).to(DEV).contiguous(memory_format=CL)   # C % 4 != 0
    assert not ops.res_tail_supported(xl, xl)


def _pair(net_name, fmt, img, classes=10, bits=4):
    from ood_dfq_b200 import fusion, nets, surgery
    torch.manual_seed(1)
    base = getattr(nets, net_name)(num_classes=classes) if net_name != "resnet18_small" else nets.resnet18_small(3, classes)
    nets.perturb_bn_stats(base)
    plain = surgery.quantize_model(base, bits, bits).to(DEV).to(memory_format=fmt).eval()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(8 if img < 100 else 4, 3, img, img, generator=g).to(DEV).contiguous(memory_format=fmt) for _ in range(3)]
    with torch.no_grad():
        for x in xs:
            plain(x)
    surgery.freeze_model(plain)
    fusion.fuse_eval_bn(plain, xs[0][:2])
    fused = copy.deepcopy(plain)
    n = fusion.fuse_residual_tails(fused, xs[0][:2])
    return plain, fused, xs, n


@pytest.mark.parametrize("net,img,units", [("resnet20_cifar", 32, 9), ("resnet18_imagenet", 224, 8), ("resnet18_small", 28, 8)])
def test_fused_tail_model_matches_unfused_model(net, img, units):
    """Whole student, channels_last: same logits, same input gradient, same parameter gradients."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    plain, fused, xs, n = _pair(net, CL, img)
    assert n == units
    a, b = xs[2].clone().requires_grad_(True), xs[2].clone().requires_grad_(True)
    ya, yb = plain(a), fused(b)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    ya.square().mean().backward()
    yb.square().mean().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-4, atol=1e-7 + 1e-5 * a.grad.abs().max().item())
    for (n1, p1), (n2, p2) in zip(plain.named_parameters(), fused.named_parameters()):
        assert n1 == n2 and torch.allclose(p1.grad, p2.grad, rtol=2e-3, atol=1e-5 * (1 + p1.grad.abs().max().item())), n1
    assert list(plain.state_dict()) == list(fused.state_dict())
    # NCHW input, training mode and a calibrating QuantAct all take the class's own forward
    with torch.no_grad():
        yc = fused(xs[2].contiguous())
    assert torch.allclose(yc, yb.detach(), rtol=1e-3, atol=1e-3)
    torch.backends.cudnn.deterministic = False


def test_fused_tail_survives_deepcopy_and_foreign_hooks():
    plain, fused, xs, _ = _pair("resnet20_cifar", CL, 32)
    clone = copy.deepcopy(fused)
    for p in clone.parameters():
        p.data.mul_(1.5)
    with torch.no_grad():
        y0, y1 = fused(xs[0]), plain(xs[0])
    assert torch.allclose(y0, y1, rtol=1e-5, atol=1e-6)          # the clone's plan points at the clone's modules
    from ood_dfq_b200 import nets
    seen = []
    unit = next(m for m in fused.modules() if isinstance(m, nets.ResUnit))
    h = unit.body.conv2.bn.register_forward_pre_hook(lambda m, i: seen.append(i[0].shape))
    with torch.no_grad():
        y2 = fused(xs[0])
    h.remove()
    assert len(seen) == 1 and torch.allclose(y2, y1, rtol=1e-5, atol=1e-6)   # hooked unit ran unfused, hook fired


def test_qat_step_with_fused_tails_matches_step_without():
    """The whole data-free QAT iteration (feature-alignment taps included): same loss, same update."""
    from ood_dfq_b200 import fusion, nets, step, surgery
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True

    def build(tails):
        torch.manual_seed(1)
        teacher = nets.resnet20_cifar(num_classes=10)
        nets.perturb_bn_stats(teacher)
        student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4).to(DEV).to(memory_format=CL)
        teacher = teacher.to(DEV).to(memory_format=CL)
        g = torch.Generator().manual_seed(2)
        xs = [torch.randn(16, 3, 32, 32, generator=g).to(DEV).contiguous(memory_format=CL) for _ in range(4)]
        with torch.no_grad():
            for x in xs[:2]:
                student(x)
        surgery.freeze_model(student)
        fusion.fuse_eval_bn(student, xs[0][:2])
        fusion.fuse_eval_bn(teacher, xs[0][:2])
        if tails:
            assert fusion.fuse_residual_tails(student, xs[0][:2]) == 9
            assert fusion.fuse_residual_tails(teacher, xs[0][:2]) == 9
        return student, step.QATStep(student, teacher, lr=1e-5, unit_types=(nets.ResUnit,)), xs

    s0, q0, xs = build(False)
    s1, q1, _ = build(True)
    for x in xs:
        l0, l1 = q0(x), q1(x)
        assert l0.item() == l0.item() and abs(l0.item() - l1.item()) <= 1e-4 * abs(l0.item()) + 1e-6
        assert len(q1.tap_s.maps) == len(q0.tap_s.maps) == 9 and len(q1.tap_t.maps) == 9
    for (n0, p0), (n1, p1) in zip(s0.named_parameters(), s1.named_parameters()):
        assert torch.allclose(p0, p1, rtol=1e-4, atol=1e-6), n0
    torch.backends.cudnn.deterministic = False


def test_input_gradient_only_sweep_skips_parameter_reductions_and_keeps_the_image_gradient():
    """``fusion.input_gradient_only()`` around ``autograd.grad(loss, images)`` (the sign perturbation of
    trainer_direct.py:508-512): the fused backwards run their gradient-only variants -- fewer launches (no fold
    kernels), same image gradient bit for bit."""
    import copy
    from ood_dfq_b200 import _native, fusion, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.manual_seed(3)
    base = nets.perturb_bn_stats(nets.resnet18_imagenet(num_classes=10))
    net = surgery.quantize_model(copy.deepcopy(base), 4, 4, namespace=qm).to(DEV).to(memory_format=torch.channels_last).eval()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 224, 224, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            net(x)
    surgery.freeze_model(net, qm)
    fusion.fuse_eval_bn(net, x[:2])
    fusion.fuse_residual_tails(net, x[:2])
    torch.backends.cudnn.deterministic = True

    def sweep(skip):
        xi = x.detach().clone().requires_grad_(True)
        loss = net(xi).square().mean()
        torch.cuda.synchronize()
        _native.reset_launch_count()
        if skip:
            with fusion.input_gradient_only():
                gx = torch.autograd.grad(loss, xi)[0]
        else:
            gx = torch.autograd.grad(loss, xi)[0]
        return gx, _native.launch_count()

    g_full, n_full = sweep(False)
    g_skip, n_skip = sweep(True)
    assert torch.equal(g_full.view(torch.int32), g_skip.view(torch.int32))
    assert n_skip < n_full, (n_skip, n_full)
    assert not fusion._SKIP_PARAM_GRADS
