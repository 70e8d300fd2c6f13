"""Fused eval-mode BatchNorm + ReLU + QuantAct against the unfused chain."""
import copy

import numpy as np
import pytest
import torch

from conftest import bits
from oracle import fq_torch, fused_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [(8, 64, 56, 56), (4, 16, 112, 112), (16, 128, 28, 28), (16, 256, 14, 14), (32, 512, 7, 7), (32, 64, 4, 4),
          (3, 5, 7, 9), (2, 3, 1, 1), (5, 130, 7, 7), (2, 6, 70, 70), (4, 2, 33, 35), (9, 16, 8, 8)]


def make_bn(c, g, affine=True):
    w = (torch.rand(c, generator=g) + 0.5) if affine else None
    b = (torch.randn(c, generator=g) * 0.3) if affine else None
    return w, b, torch.randn(c, generator=g) * 0.2, torch.rand(c, generator=g) + 0.4


def cu(t):
    return None if t is None else t.to(DEV)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("relu,k", [(True, 4), (True, 2), (True, 0), (False, 0), (False, 8)])
def test_fused_forward(shape, relu, k):
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + 7 * k + relu)
    x = torch.randn(shape, generator=g) * 1.4
    w, b, rm, rv = make_bn(shape[1], g, affine=(k != 2))
    eps = 1e-5
    lo, hi = torch.tensor([0.0 if relu else -1.2]), torch.tensor([1.9])
    fq = (k, cu(lo), cu(hi)) if k else None
    y, z = ops.bn_eval_forward(cu(x), cu(w), cu(b), cu(rm), cu(rv), eps, relu=relu, fq=fq, want_z=True)
    z64 = fused_torch.bn_eval_affine64(x, w, b, rm, rv, eps)
    if relu:
        z64 = z64.clamp_min(0)
    # the affine is a few fp32 roundings away from the exact value (a_c, b_c are rounded, then one FMA)
    np.testing.assert_allclose(z.cpu().double().numpy(), z64.numpy(), rtol=2e-6, atol=2e-6)
    if k:   # the quantiser applied to THAT fp32 value is the reference's arithmetic, bit for bit
        ref = fq_torch.fake_quant(z.cpu(), k, lo, hi)
        assert np.array_equal(bits(y.cpu().numpy()), bits(ref.numpy()))
    else:
        assert torch.equal(y, z)
    y2 = ops.bn_eval_forward(cu(x), cu(w), cu(b), cu(rm), cu(rv), eps, relu=relu, fq=fq)     # no debug output
    assert torch.equal(y2, y)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("relu", [True, False])
def test_fused_backward(shape, relu):
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + relu)
    x = torch.randn(shape, generator=g) * 1.4
    w, b, rm, rv = make_bn(shape[1], g)
    gy = torch.randn(shape, generator=g)
    eps = 1e-5
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = fused_torch.bn_relu_quant(xr, wr, br, rm, rv, eps, 4, torch.zeros(1), torch.ones(1) * 2, relu=relu)
    out.backward(gy)
    gx, dw, db = ops.bn_eval_backward(cu(x), cu(gy), cu(w), cu(b), cu(rm), cu(rv), eps, relu=relu)
    # the ReLU mask is recomputed from x: elements within rounding of 0 may differ from the CPU chain
    diff = (gx.cpu() - xr.grad).abs()
    tol = 1e-5 * xr.grad.abs().max().item()
    assert (diff > tol).float().mean().item() < 1e-5
    n_red = x.numel() / shape[1]
    np.testing.assert_allclose(dw.cpu().numpy(), wr.grad.numpy(), rtol=2e-4, atol=2e-5 * n_red ** 0.5)
    np.testing.assert_allclose(db.cpu().numpy(), br.grad.numpy(), rtol=2e-4, atol=2e-5 * n_red ** 0.5)
    gx2, dw2, db2 = ops.bn_eval_backward(cu(x), cu(gy), cu(w), cu(b), cu(rm), cu(rv), eps, relu=relu,
                                         want_param_grads=False)
    assert dw2 is None and torch.equal(gx2, gx)


@pytest.mark.parametrize("net_name,classes,k,shape", [
    ("resnet20_cifar", 10, 4, (32, 3, 32, 32)),
    ("resnet18_small", 9, 2, (16, 3, 28, 28)),
    ("resnet18_imagenet", 1000, 4, (4, 3, 224, 224)),
])
def test_fusion_pass_matches_unfused_model(net_name, classes, k, shape):
    """Same student, fused vs unfused, on the GPU: logits, input gradient and every parameter gradient."""
    from ood_dfq_b200 import fusion, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    base = nets.resnet18_small(3, classes) if net_name == "resnet18_small" else getattr(nets, net_name)(num_classes=classes)
    nets.perturb_bn_stats(base)
    plain = surgery.quantize_model(base, k, k).to(DEV).eval()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(shape, generator=g).to(DEV) for _ in range(3)]
    with torch.no_grad():
        for x in xs:                                  # calibrate, then freeze
            plain(x)
    surgery.freeze_model(plain)
    fused = copy.deepcopy(plain)
    keys = list(fused.state_dict())
    fusion.fuse_eval_bn(fused, xs[0][:2])
    assert list(fused.state_dict()) == keys
    n_bn = sum(isinstance(m, torch.nn.BatchNorm2d) for m in plain.modules())
    assert sum(type(m) is fusion.FusedEvalBN for m in fused.modules()) == n_bn
    assert sum(type(m) is fusion.AbsorbedTail for m in fused.modules()) >= 9
    x = xs[2].clone().requires_grad_(True)
    x2 = xs[2].clone().requires_grad_(True)
    yp, yf = plain(x), fused(x2)
    # cuDNN's inference BN and the fused affine round differently: a handful of codes move by one step
    spread = yp.std().item()
    assert (yf - yp).abs().max().item() < 0.2 * spread
    yp.square().mean().backward()
    yf.square().mean().backward()
    cos = torch.nn.functional.cosine_similarity(x.grad.flatten(), x2.grad.flatten(), dim=0).item()
    assert cos > 0.98, cos
    for (n1, p1), (n2, p2) in zip(plain.named_parameters(), fused.named_parameters()):
        assert n1 == n2 and p2.grad is not None, n1
        n_a, n_b = p1.grad.norm().item(), p2.grad.norm().item()
        if n_a == 0.0 and n_b == 0.0:
            continue                                  # dead layer in this random-init network: both agree on zero
        c = torch.nn.functional.cosine_similarity(p1.grad.flatten(), p2.grad.flatten(), dim=0).item()
        assert c > 0.95 and 0.8 < n_b / n_a < 1.25, (n1, c, n_a, n_b)
    # calibrating again falls back to the exact unfused chain and tracks ranges identically
    surgery.unfreeze_model(plain)
    surgery.unfreeze_model(fused)
    with torch.no_grad():
        plain(xs[1])
        fused(xs[1])
    for a, b in zip([m for m in plain.modules() if type(m) is qm.QuantAct],
                    [m for m in fused.modules() if type(m) is qm.QuantAct]):
        assert torch.equal(a.x_max, b.x_max) and torch.equal(a.beta_t, b.beta_t)


NHWC_SHAPES = [(8, 64, 56, 56), (16, 128, 28, 28), (32, 512, 7, 7), (4, 24, 9, 11), (3, 960, 5, 5), (2, 1280, 3, 3),
               (5, 8, 1, 7)]


@pytest.mark.parametrize("shape", NHWC_SHAPES)
@pytest.mark.parametrize("relu,k", [(True, 4), (False, 0), (True, 0)])
def test_fused_channels_last_matches_nchw(shape, relu, k):
    """The NHWC kernels compute exactly what the NCHW kernels compute (same arithmetic per element)."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = (torch.randn(shape, generator=g) * 1.4).to(DEV)
    gy = torch.randn(shape, generator=g).to(DEV)
    w, b, rm, rv = (cu(t) for t in make_bn(shape[1], g))
    lo, hi = torch.zeros(1, device=DEV), torch.full((1,), 1.9, device=DEV)
    fq = (k, lo, hi) if k else None
    y = ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=relu, fq=fq)
    xl, gl = x.contiguous(memory_format=torch.channels_last), gy.contiguous(memory_format=torch.channels_last)
    yl = ops.bn_eval_forward(xl, w, b, rm, rv, 1e-5, relu=relu, fq=fq)
    assert yl.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(yl, y)
    gx, dw, db = ops.bn_eval_backward(x, gy, w, b, rm, rv, 1e-5, relu=relu)
    gxl, dwl, dbl = ops.bn_eval_backward(xl, gl, w, b, rm, rv, 1e-5, relu=relu)
    assert torch.equal(gxl, gx)
    n_red = x.numel() / shape[1]
    np.testing.assert_allclose(dwl.cpu().numpy(), dw.cpu().numpy(), rtol=1e-4, atol=1e-5 * n_red ** 0.5)
    np.testing.assert_allclose(dbl.cpu().numpy(), db.cpu().numpy(), rtol=1e-4, atol=1e-5 * n_red ** 0.5)


def test_channels_last_weights_and_activations():
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(8)
    w = (torch.randn(32, 16, 3, 3, generator=g) * 0.05).to(DEV)
    wl = w.contiguous(memory_format=torch.channels_last)
    (r,) = ops.weight_fq_multi([w], [4], [False], want_range=True, want_codes=True)
    (rl,) = ops.weight_fq_multi([wl], [4], [False], want_range=True, want_codes=True)
    assert rl["wq"].is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(rl["wq"], r["wq"]) and torch.equal(rl["codes"], r["codes"]) and torch.equal(rl["lo"], r["lo"])
    ref = fq_torch.fake_quant(w.cpu(), 4, *fq_torch.row_minmax(w.cpu()))
    assert np.array_equal(bits(rl["wq"].cpu().contiguous().numpy()), bits(ref.numpy()))


def test_channels_last_student_matches_nchw_student():
    """Whole fused student in channels_last vs NCHW: same logits / gradients up to convolution rounding."""
    from ood_dfq_b200 import fusion, nets, step, surgery
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    base = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(base)
    a = surgery.quantize_model(base, 4, 4).to(DEV).eval()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(16, 3, 32, 32, generator=g).to(DEV) for _ in range(3)]
    with torch.no_grad():
        for x in xs:
            a(x)
    surgery.freeze_model(a)
    b = copy.deepcopy(a).to(memory_format=torch.channels_last)
    fusion.fuse_eval_bn(a, xs[0][:2])
    fusion.fuse_eval_bn(b, xs[0][:2].contiguous(memory_format=torch.channels_last))
    xa = xs[2].clone().requires_grad_(True)
    xb = xs[2].clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ya, yb = a(xa), b(xb)
    assert (ya - yb).abs().max().item() < 0.2 * ya.std().item()
    ya.square().mean().backward()
    yb.square().mean().backward()
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        n_a, n_b = p1.grad.norm().item(), p2.grad.norm().item()
        if n_a == 0.0 and n_b == 0.0:
            continue
        c = torch.nn.functional.cosine_similarity(p1.grad.flatten(), p2.grad.flatten(), dim=0).item()
        assert c > 0.95 and 0.8 < n_b / n_a < 1.25, (n1, c, n_a, n_b)
    # flat gradients keep the parameters' layout and notice a later .to()
    teacher = copy.deepcopy(base).to(DEV).to(memory_format=torch.channels_last)
    qat = step.QATStep(b, teacher, unit_types=(nets.ResUnit,))
    assert all(p.grad.stride() == p.stride() for p in qat.grads.params)
    qat(xs[0].contiguous(memory_format=torch.channels_last))
    b.to(memory_format=torch.contiguous_format)
    with pytest.raises(RuntimeError, match="flat buffer"):
        qat(xs[0])


@pytest.mark.parametrize("all_passes", [False, True])
def test_graphed_step_matches_eager_step(all_passes):
    """The whole QAT iteration replayed as a CUDA graph updates the student exactly like the eager step
    (``all_passes``: channels_last with the residual-tail fusion and the twin handles as well)."""
    from ood_dfq_b200 import fusion, nets, step, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True

    def build():
        torch.manual_seed(1)
        teacher = nets.resnet20_cifar(num_classes=10)
        nets.perturb_bn_stats(teacher)
        fmt = torch.channels_last if all_passes else torch.contiguous_format
        student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4).to(DEV).to(memory_format=fmt)
        teacher = teacher.to(DEV).to(memory_format=fmt)
        g = torch.Generator().manual_seed(2)
        xs = [torch.randn(16, 3, 32, 32, generator=g).to(DEV).contiguous(memory_format=fmt) for _ in range(4)]
        with torch.no_grad():
            for x in xs[:2]:
                student(x)
        surgery.freeze_model(student)
        fusion.fuse_eval_bn(student, xs[0][:2])
        if all_passes:
            fusion.fuse_eval_bn(teacher, xs[0][:2])
            assert fusion.fuse_residual_tails(student, xs[0][:2]) == 9
            assert fusion.fuse_residual_tails(teacher, xs[0][:2]) == 9
        qat = step.QATStep(student, teacher, lr=1e-5, unit_types=(nets.ResUnit,))
        return student, qat, xs

    s_eager, q_eager, xs = build()
    s_graph, q_graph, _ = build()
    graphed = step.GraphedStep(q_graph, xs[0], warmup=2)          # 2 eager warm-up updates; capture itself runs nothing
    for _ in range(2):
        q_eager(xs[0])
    losses = []
    for x in xs[1:]:
        le, lg = q_eager(x), graphed(x)
        losses.append((le.item(), lg.item()))
    for le, lg in losses:
        assert le == le and abs(le - lg) <= 1e-4 * abs(le) + 1e-6, losses
    for (n1, p1), (n2, p2) in zip(s_eager.named_parameters(), s_graph.named_parameters()):
        assert torch.allclose(p1, p2, rtol=1e-4, atol=1e-6), n1
    # an eager forward after the replays sees the up-to-date quantised weights
    with torch.no_grad():
        ye, yg = s_eager(xs[0]), s_graph(xs[0])
    assert torch.allclose(ye, yg, rtol=1e-3, atol=1e-4)
    torch.backends.cudnn.deterministic = False


@pytest.mark.parametrize("shape", NHWC_SHAPES)
def test_bn_statistics_channels_last_matches_nchw(shape):
    from ood_dfq_b200 import bns, ops
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 1.3 + 0.4).to(DEV)
    gy = torch.randn(shape, generator=g).to(DEV)
    c = shape[1]
    rm = (torch.randn(c, generator=g) * 0.2).to(DEV)
    xl, gl = x.contiguous(memory_format=torch.channels_last), gy.contiguous(memory_format=torch.channels_last)
    lo, hi = torch.zeros(1, device=DEV), torch.full((1,), 1.7, device=DEV)
    s_a, y_a = ops.bn_stats_forward(x, rm, fq=(4, lo, hi))
    s_b, y_b = ops.bn_stats_forward(xl, rm, fq=(4, lo, hi))
    assert y_b.is_contiguous(memory_format=torch.channels_last) and torch.equal(y_a, y_b)
    np.testing.assert_allclose(s_b.cpu().numpy(), s_a.cpu().numpy(), rtol=1e-6, atol=1e-6 * x.numel() / c)
    np.testing.assert_allclose(ops.bn_stats_forward(xl, rm).cpu().numpy(), s_b.cpu().numpy(), rtol=1e-12)
    cnt = float(x.numel() // c)
    mean, var = ops.bn_stats_finalize(s_b, rm, cnt)
    np.testing.assert_allclose(mean.cpu().numpy(), x.mean([0, 2, 3]).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var.cpu().numpy(), x.var([0, 2, 3], unbiased=False).cpu().numpy(), rtol=2e-5)
    gm, gv = torch.randn(c, generator=g).to(DEV), torch.randn(c, generator=g).to(DEV)
    for gin_a, gin_b in ((gy, gl), (None, None)):
        ga = ops.bn_stats_backward(x, gin_a, mean, gm, gv, cnt)
        gb = ops.bn_stats_backward(xl, gin_b, mean, gm, gv, cnt)
        assert gb.is_contiguous(memory_format=torch.channels_last) and torch.equal(ga, gb)
    # the differentiable front end on a channels_last tensor
    xr = xl.clone().requires_grad_(True)
    m2, v2 = bns.bn_channel_stats(xr, rm)
    (m2.square().sum() + v2.square().sum()).backward()
    xn = x.clone().requires_grad_(True)
    m1, v1 = bns.bn_channel_stats(xn, rm)
    (m1.square().sum() + v1.square().sum()).backward()
    np.testing.assert_allclose(xr.grad.cpu().numpy(), xn.grad.cpu().numpy(), rtol=1e-4, atol=1e-7)


def test_fusion_handles_sync_batchnorm():
    """main_direct.py:483 converts the student to SyncBatchNorm; in eval() that is the same affine."""
    from ood_dfq_b200 import fusion, nets, surgery
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    base = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(base)
    plain = surgery.quantize_model(base, 4, 4)
    plain = torch.nn.SyncBatchNorm.convert_sync_batchnorm(plain).to(DEV).eval()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(8, 3, 32, 32, generator=g).to(DEV) for _ in range(3)]
    with torch.no_grad():
        for x in xs:
            plain(x)
    surgery.freeze_model(plain)
    fused = copy.deepcopy(plain)
    fusion.fuse_eval_bn(fused, xs[0][:2])
    assert sum(type(m) is fusion.FusedEvalSyncBN for m in fused.modules()) == 21
    assert sum(type(m) is fusion.AbsorbedTail for m in fused.modules()) == 10
    assert all(isinstance(m, torch.nn.SyncBatchNorm) for m in fused.modules() if type(m) is fusion.FusedEvalSyncBN)
    with torch.no_grad():
        yp, yf = plain(xs[2]), fused(xs[2])
    assert (yf - yp).abs().max().item() < 0.2 * yp.std().item()


def test_relu_first_flat_kernel_and_fused_tail():
    from ood_dfq_b200 import fusion, ops
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 7, 9, 11, generator=g) * 2
    x[0, 0, 0, 0] = float("nan")
    lo, hi = torch.zeros(1), torch.full((1,), 2.2)
    for k in (2, 4, 8):
        y = ops.fake_quant(x.to(DEV), k, lo.to(DEV), hi.to(DEV), relu_first=True).cpu()
        ref = fq_torch.fake_quant(torch.relu(x), k, lo, hi)
        assert torch.isnan(y[0, 0, 0, 0]) and torch.isnan(ref[0, 0, 0, 0])       # NaN payloads may differ
        assert np.array_equal(bits(torch.nan_to_num(y, 9.0).numpy()), bits(torch.nan_to_num(ref, 9.0).numpy()))
    tail = torch.nn.Sequential(torch.nn.ReLU(inplace=True), qm.QuantAct(4)).to(DEV)
    tail(torch.relu(x[1:]).to(DEV))                      # one calibrating pass
    tail[1].fix()
    plain_out_in = x[1:].clone().to(DEV).requires_grad_(True)
    yp = tail(plain_out_in * 1.0)
    tail.__class__ = fusion.FusedReLUQuant
    fused_in = x[1:].clone().to(DEV).requires_grad_(True)
    yf = tail(fused_in * 1.0)
    assert torch.equal(yp, yf)
    gy = torch.randn_like(yp)
    yp.backward(gy)
    yf.backward(gy)
    assert torch.equal(plain_out_in.grad, fused_in.grad)


def test_fused_full_precision_teacher_matches_plain():
    from ood_dfq_b200 import fusion, nets
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    plain = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(plain)
    plain = plain.to(DEV).eval()
    fused = copy.deepcopy(plain)
    x = torch.randn(8, 3, 32, 32, device=DEV)
    fusion.fuse_eval_bn(fused, x[:2])
    assert sum(type(m) is fusion.AbsorbedReLU for m in fused.modules()) == 10
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = plain(a), fused(b)
    assert torch.allclose(ya, yb, rtol=1e-4, atol=1e-4)
    ya.square().mean().backward()
    yb.square().mean().backward()
    assert torch.allclose(a.grad, b.grad, rtol=1e-3, atol=1e-6)
    for (n1, p1), (n2, p2) in zip(plain.named_parameters(), fused.named_parameters()):
        assert torch.allclose(p1.grad, p2.grad, rtol=2e-3, atol=1e-5), n1


@pytest.mark.parametrize("shape", [(8, 64, 56, 56), (16, 128, 28, 28), (32, 512, 7, 7), (3, 5, 7, 9), (4, 24, 9, 11),
                                   (2, 1280, 3, 3), (5, 8, 1, 7), (2, 16, 40, 40)])
def test_channel_energy_matches_reference_expression(shape):
    """The feature-alignment reduction (trainer_direct.py:382-383), NCHW and channels_last, forward and backward."""
    from ood_dfq_b200 import step
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 1.3).to(DEV)
    for fmt in (torch.contiguous_format, torch.channels_last):
        a = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        b = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        ref, out = step.channel_attention(a.clone()), step.channel_attention_fused(b)
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-5, atol=1e-7)
        w = torch.randn_like(ref)
        (ref * w).sum().backward()
        (out * w).sum().backward()
        scale = a.grad.abs().max().item()
        np.testing.assert_allclose(b.grad.cpu().numpy(), a.grad.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)
        assert b.grad.stride() == b.stride()


@pytest.mark.parametrize("register_kernel", [False, True])
@pytest.mark.parametrize("shape", [(4, 64, 112, 112), (3, 16, 9, 11), (2, 8, 7, 7), (2, 128, 5, 6), (1, 4, 1, 1)])
@pytest.mark.parametrize("k", [4, 2, 0])
def test_fused_stem_matches_unfused_chain(shape, k, register_kernel):
    """BN -> ReLU -> [QuantAct] -> MaxPool2d(3,2,1) as one kernel vs the fused BN kernel followed by ATen's
    max_pool2d: same values, same argmax tie-breaking (first maximum), hence the same gradient field.  Both
    forward kernels: the TMA-staged ring (default) and the register-staged fallback."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = (torch.randn(shape, generator=g) * 1.4).to(DEV).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = (cu(t) for t in make_bn(shape[1], g))
    lo, hi = torch.zeros(1, device=DEV), torch.full((1,), 1.9, device=DEV)
    fq = (k, lo, hi) if k else None
    y = ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=True, fq=fq).requires_grad_(True)
    ref = torch.nn.functional.max_pool2d(y, 3, 2, 1)
    out, idx, xhat = ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, register_kernel=register_kernel)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, ref)
    go = torch.randn(ref.shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    ref.backward(go)
    gx_ref, dw_ref, db_ref = ops.bn_eval_backward(x, y.grad, w, b, rm, rv, 1e-5, relu=True)
    gx, dw, db = ops.bn_pool_backward(go, idx, xhat, x.shape, w, b, rm, rv, 1e-5)
    assert torch.equal(gx, gx_ref)
    n_red = x.numel() / shape[1]
    np.testing.assert_allclose(dw.cpu().numpy(), dw_ref.cpu().numpy(), rtol=2e-4, atol=2e-5 * n_red ** 0.5)
    np.testing.assert_allclose(db.cpu().numpy(), db_ref.cpu().numpy(), rtol=2e-4, atol=2e-5 * n_red ** 0.5)
    gx2, dw2, _ = ops.bn_pool_backward(go, idx, None, x.shape, w, b, rm, rv, 1e-5, want_param_grads=False)
    assert dw2 is None and torch.equal(gx2, gx)


@pytest.mark.parametrize("register_kernel", [False, True])
@pytest.mark.parametrize("case", ["merged_table", "nan", "odd_sizes_k2", "tall_segments"])
def test_fused_stem_corner_cases(case, register_kernel):
    """Degenerate dequantisation table (several codes share one value, so ties must be broken on the VALUE),
    NaN inputs (the last NaN of a window wins, as in ATen), odd extents, and a batch small enough that every
    window column is split into several row segments."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(7)
    k, lo, hi, shape = 4, 0.0, 1.9, (3, 16, 12, 10)
    if case == "odd_sizes_k2":
        k, shape = 2, (2, 12, 13, 7)
    if case == "tall_segments":
        shape = (1, 8, 90, 34)
    x = torch.randn(shape, generator=g) * 1.4
    if case == "nan":
        x.view(-1)[torch.randperm(x.numel(), generator=g)[:40]] = float("nan")
    x = x.to(DEV).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = (cu(t) for t in make_bn(shape[1], g))
    if case == "merged_table":
        # range sitting at 2^23: dequantised values have ulp 1 while codes are 4/15 apart -> merged table entries
        lo, hi = float(2 ** 23), float(2 ** 23 + 4)
        b = b + float(2 ** 23 + 2)
    lo_t, hi_t = torch.full((1,), lo, device=DEV), torch.full((1,), hi, device=DEV)
    fq = (k, lo_t, hi_t)
    y = ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=True, fq=fq).requires_grad_(True)
    if case == "merged_table":
        assert y.detach().unique().numel() < 2 ** k          # the table really has merged entries
    ref = torch.nn.functional.max_pool2d(y, 3, 2, 1)
    out, idx, xhat = ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, register_kernel=register_kernel)
    assert np.array_equal(out.cpu().numpy().view(np.int32), ref.detach().cpu().numpy().view(np.int32)) or \
        (case == "nan" and torch.equal(torch.isnan(out), torch.isnan(ref)) and
         torch.equal(torch.nan_to_num(out), torch.nan_to_num(ref.detach())))
    go = torch.randn(ref.shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    ref.backward(go)
    gx_ref, _, _ = ops.bn_eval_backward(x, y.grad, w, b, rm, rv, 1e-5, relu=True)
    gx, _, _ = ops.bn_pool_backward(go, idx, xhat, x.shape, w, b, rm, rv, 1e-5)
    if case == "nan":
        # the ReLU mask of a NaN activation is implementation-defined in both chains; compare elsewhere
        keep = ~torch.isnan(x)
        assert torch.equal(gx[keep], gx_ref[keep])
    else:
        assert torch.equal(gx, gx_ref)


@pytest.mark.parametrize("shape", [(2, 64, 112, 112), (5, 64, 56, 56), (3, 32, 33, 47), (2, 16, 14, 5), (7, 8, 2, 2),
                                   (1, 12, 3, 1), (2, 256, 9, 16), (300, 16, 10, 10), (2, 64, 224, 224)])
@pytest.mark.parametrize("mode", ["k4", "k8", "k1", "plain", "merged", "nan", "ties"])
def test_stem_ring_kernel_equals_register_kernel(shape, mode):
    """The TMA-staged stem forward (packed integer candidates, winner lookups in the ring) against the register
    kernel: output bits, argmax / ReLU bytes and the normalised input at the argmax are identical -- many items per
    CTA, rows that wrap the ring, odd extents, one-window rows, all-equal inputs (every tie), NaNs (the last NaN of
    a window wins in both) and a table with merged entries (candidates compared on the value)."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + len(mode))
    x = torch.randn(shape, generator=g) * 1.4
    if mode == "ties":
        x = torch.round(x)                       # few distinct values: ties inside most windows
        x[0] = 0.25                              # and an image where every window is one big tie
    if mode == "nan":
        x.view(-1)[torch.randperm(x.numel(), generator=g)[:max(3, x.numel() // 50)]] = float("nan")
    x = x.to(DEV).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = (cu(t) for t in make_bn(shape[1], g))
    lo, hi, k = 0.0, 1.9, {"k8": 8, "k1": 1}.get(mode, 4)
    if mode == "merged":
        lo, hi = float(2 ** 23), float(2 ** 23 + 4)
        b = b + float(2 ** 23 + 2)
    fq = None if mode == "plain" else (k, torch.full((1,), lo, device=DEV), torch.full((1,), hi, device=DEV))
    for want_xhat in (True, False):
        o1, i1, x1 = ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, want_xhat=want_xhat)
        o2, i2, x2 = ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, want_xhat=want_xhat, register_kernel=True)
        assert torch.equal(o1.view(torch.int32), o2.view(torch.int32))
        assert torch.equal(i1, i2)
        if want_xhat:
            assert torch.equal(x1.view(torch.int32), x2.view(torch.int32))
        else:
            assert x1 is None and x2 is None


def test_fused_stem_in_the_imagenet_student():
    from ood_dfq_b200 import fusion, nets, surgery
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    base = nets.resnet18_imagenet(num_classes=10)
    nets.perturb_bn_stats(base)
    plain = surgery.quantize_model(base, 4, 4).to(DEV).to(memory_format=torch.channels_last).eval()
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(4, 3, 224, 224, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) for _ in range(3)]
    with torch.no_grad():
        for x in xs:
            plain(x)
    surgery.freeze_model(plain)
    fused = copy.deepcopy(plain)
    fusion.fuse_eval_bn(fused, xs[0][:2])
    assert sum(type(m) is fusion.AbsorbedPool for m in fused.modules()) == 1
    a, b = xs[2].clone().requires_grad_(True), xs[2].clone().requires_grad_(True)
    ya, yb = plain(a), fused(b)
    assert (ya - yb).abs().max().item() < 0.2 * ya.std().item()
    ya.square().mean().backward()
    yb.square().mean().backward()
    assert torch.nn.functional.cosine_similarity(a.grad.flatten(), b.grad.flatten(), dim=0).item() > 0.98
    stem_p = dict(plain.named_parameters())["features.0.conv.bn.weight"].grad
    stem_f = dict(fused.named_parameters())["features.0.conv.bn.weight"].grad
    assert stem_f is not None and torch.nn.functional.cosine_similarity(stem_p, stem_f, dim=0).item() > 0.98
    # NCHW input falls back to fused-BN + ordinary pooling
    with torch.no_grad():
        yc = fused(xs[2].contiguous())
    assert torch.allclose(yc, yb.detach(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("shape", [(4, 64, 12, 12), (32, 128, 28, 28), (3, 8, 5, 7), (2, 512, 7, 7)])
@pytest.mark.parametrize("k", [4, 0])
def test_fused_bn_relu_mask_replaces_x_in_the_gradient_only_backward(shape, k):
    """channels_last + ReLU: the forward can leave one byte per four channels saying where the ReLU is open; a
    backward that wants no parameter gradients then takes the mask INSTEAD of x -- same bits as the pass that
    re-derives the mask from x (NaN closes the ReLU in both)."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = (torch.randn(shape, generator=g) * 1.4)
    x[0, 0, 0, 0] = float("nan")
    x = cu(x).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = (cu(t) for t in make_bn(shape[1], g))
    fq = (k, cu(torch.zeros(1)), cu(torch.tensor([1.9]))) if k else None
    y, mask = ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=True, fq=fq, want_mask=True)
    assert torch.equal(y.view(torch.int32), ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=True, fq=fq).view(torch.int32))
    assert mask.dtype == torch.uint8 and mask.numel() == x.numel() // 4
    gy = cu(torch.randn(shape, generator=g)).contiguous(memory_format=torch.channels_last)
    ref, _, _ = ops.bn_eval_backward(x, gy, w, b, rm, rv, 1e-5, relu=True, want_param_grads=False)
    got, dw, db = ops.bn_eval_backward(None, gy, w, b, rm, rv, 1e-5, relu=True, want_param_grads=False, mask=mask)
    assert dw is None and db is None and got.stride() == ref.stride()
    assert torch.equal(got.view(torch.int32), ref.view(torch.int32))
    # with parameter gradients the mask is simply not used (x is needed for dW anyway)
    full = ops.bn_eval_backward(x, gy, w, b, rm, rv, 1e-5, relu=True, mask=mask)
    base = ops.bn_eval_backward(x, gy, w, b, rm, rv, 1e-5, relu=True)
    assert all(torch.equal(a.view(torch.int32), b_.view(torch.int32)) for a, b_ in zip(full, base))     # (NaN-safe)
    # NCHW tensors have no mask variant: the forward says so by returning None
    _, none = ops.bn_eval_forward(x.contiguous(), w, b, rm, rv, 1e-5, relu=True, fq=fq, want_mask=True)
    assert none is None
    with pytest.raises(RuntimeError, match="needs x"):
        ops.bn_eval_backward(None, gy, w, b, rm, rv, 1e-5, relu=True)


@pytest.mark.parametrize("quantised", [False, True])
def test_statistics_tap_and_fused_batchnorm_share_one_backward_pass(quantised):
    """``bns.BNStatLoss`` on a network whose BatchNorms are fused (the distillation loop hooks EVERY BatchNorm,
    distill_data.py:69-78): tap and BatchNorm form one autograd node whose backward is a single kernel.  Loss and image
    gradient must equal the two-kernel chain (a fused BatchNorm with the ordinary pre-hook tap) bit for bit."""
    from ood_dfq_b200 import _native, bns, fusion, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.manual_seed(5)
    base = nets.perturb_bn_stats(nets.resnet20_cifar(num_classes=10))
    if quantised:
        base = surgery.quantize_model(base, 4, 4, namespace=qm)
    net = base.to(DEV).to(memory_format=torch.channels_last).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(8, 3, 32, 32, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    if quantised:
        with torch.no_grad():
            for _ in range(2):
                net(x)
        surgery.freeze_model(net, qm)
    fusion.fuse_eval_bn(net, x[:2])

    def run(one_node):
        fusion._FusedEvalMixin._oodfq_accepts_tap = one_node
        try:
            stat = bns.BNStatLoss(net)
            xi = x.detach().clone().requires_grad_(True)
            out = net(xi)
            loss = stat.loss() + out.square().mean()
            torch.cuda.synchronize()
            _native.reset_launch_count()
            gx = torch.autograd.grad(loss, xi)[0]
            n = _native.launch_count()
            stat.remove()
            return loss.detach(), gx, n
        finally:
            fusion._FusedEvalMixin._oodfq_accepts_tap = True

    l_chain, g_chain, n_chain = run(False)
    l_one, g_one, n_one = run(True)
    # the statistics come out of another kernel (grid, hence fp32 partial grouping, differs): equal to fp64 accuracy
    np.testing.assert_allclose(l_one.item(), l_chain.item(), rtol=1e-6)
    scale = g_chain.abs().max().item()
    np.testing.assert_allclose(g_one.cpu().numpy(), g_chain.cpu().numpy(), rtol=1e-4, atol=1e-6 * scale)
    assert n_one < n_chain, (n_one, n_chain)


@pytest.mark.parametrize("shape", [(8, 64, 14, 14), (32, 128, 28, 28), (3, 8, 5, 7), (2, 512, 7, 7)])
@pytest.mark.parametrize("relu,k", [(True, 4), (True, 0), (False, 0)])
def test_tapped_batchnorm_kernels_equal_their_two_kernel_chains(shape, relu, k):
    """Op level.  Forward: statistics of x and the fused BatchNorm output from one read -- y bit-identical to
    ``bn_eval_forward``, sums equal to ``bn_stats_forward`` up to the fp32 grouping of another grid.  Backward: BN backward
    and the BN-statistics loss gradient in one pass -- bit-identical to ``bn_eval_backward`` followed by
    ``bn_stats_backward`` accumulating into its result."""
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + k + relu)
    c = shape[1]
    x = cu(torch.randn(shape, generator=g) * 1.4 + 0.3).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = (cu(t) for t in make_bn(c, g))
    fq = (k, cu(torch.zeros(1)), cu(torch.tensor([1.9]))) if k else None
    sums = torch.empty(2 * c, dtype=torch.float64, device=DEV)
    y = ops.bn_eval_stats_forward(x, w, b, rm, rv, 1e-5, rm, sums, relu=relu, fq=fq)
    y_ref = ops.bn_eval_forward(x, w, b, rm, rv, 1e-5, relu=relu, fq=fq)
    assert y.stride() == y_ref.stride() and torch.equal(y.view(torch.int32), y_ref.view(torch.int32))
    sums_ref = ops.bn_stats_forward(x, rm)
    np.testing.assert_allclose(sums.cpu().numpy(), sums_ref.cpu().numpy(), rtol=1e-6, atol=1e-6 * x.numel() / c)
    count = float(x.numel() // c)
    mean, var = ops.bn_stats_finalize(sums_ref, rm, count)
    gmean, gvar = cu(torch.randn(c, generator=g)), cu(torch.randn(c, generator=g))
    gy = cu(torch.randn(shape, generator=g)).contiguous(memory_format=torch.channels_last)
    gs = cu(torch.tensor([0.37]))
    chain, _, _ = ops.bn_eval_backward(x, gy, w, b, rm, rv, 1e-5, relu=relu, want_param_grads=False)
    chain = ops.bn_stats_backward(x, chain, mean, gmean, gvar, count, gscale=gs)
    one = ops.bn_eval_tap_backward(x, gy, w, b, rm, rv, 1e-5, mean, gmean, gvar, count, relu=relu, gscale=gs)
    assert torch.equal(one.view(torch.int32), chain.view(torch.int32))


@pytest.mark.parametrize("net_name", ["resnet20_cifar", "resnet18_small"])
def test_fused_residual_tails_keep_running_under_the_statistics_taps(net_name):
    """The distillation loop hooks EVERY BatchNorm (distill_data.py:69-78), including the two a fused residual tail
    bypasses.  The fused unit runs those taps itself on the tensors the BatchNorms would have received: loss, per-layer
    statistics and the image gradient equal those of the network whose units run their modules one by one."""
    from ood_dfq_b200 import _native, bns, fusion, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.manual_seed(11)
    base = nets.resnet18_small(3, 9) if net_name == "resnet18_small" else nets.resnet20_cifar(num_classes=10)
    base = surgery.quantize_model(nets.perturb_bn_stats(base), 4, 4, namespace=qm)
    net = base.to(DEV).to(memory_format=torch.channels_last).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(12)
    side = 28 if net_name == "resnet18_small" else 32
    x = torch.randn(8, 3, side, side, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            net(x)
    surgery.freeze_model(net, qm)
    fusion.fuse_eval_bn(net, x[:2])
    plain = copy.deepcopy(net)
    units = fusion.fuse_residual_tails(net, x[:2])
    assert units >= 8

    def run(model):
        stat = bns.BNStatLoss(model)
        xi = x.detach().clone().requires_grad_(True)
        out = model(xi)
        loss = stat.loss() + out.square().mean()
        means = [m.clone() for m in stat.means()]
        _native.reset_launch_count()
        gx = torch.autograd.grad(loss, xi)[0]
        n = _native.launch_count()
        stat.remove()
        return loss.detach(), means, gx, n

    l_ref, m_ref, g_ref, n_ref = run(plain)
    l_new, m_new, g_new, n_new = run(net)
    np.testing.assert_allclose(l_new.item(), l_ref.item(), rtol=1e-5)
    for a, b in zip(m_new, m_ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-6)
    scale = g_ref.abs().max().item()
    # the fused tail rounds where the chain of fused BatchNorm, add and quantiser rounds (csrc/res_tail.cu), so
    # activations agree bit for bit and the gradient differs only through the fp32 grouping of the statistics sums
    np.testing.assert_allclose(g_new.cpu().numpy(), g_ref.cpu().numpy(), rtol=1e-4, atol=1e-6 * scale)
    # and the units really took the fused path (the library then launches the tail kernels: its own count goes UP,
    # what disappears are ATen's residual adds, ReLU masks and gradient-accumulation adds)
    assert any(isinstance(m, fusion._FusedUnitMixin) for m in net.modules())
    assert n_new != n_ref, (n_new, n_ref)
