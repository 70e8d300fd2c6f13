"""Host logic of the fusion passes on CPU: the class swaps keep the module tree, the state_dict and the results
(every fused module falls back to its class's own forward off the GPU)."""
import copy

import torch

from ood_dfq_b200 import fusion, nets


def _student(name):
    torch.manual_seed(1)
    base = nets.resnet18_small(3, 9) if name == "resnet18_small" else getattr(nets, name)(num_classes=10)
    nets.perturb_bn_stats(base)
    # the full-precision teacher: the quantised student's modules have no CPU path at all (by design), the
    # passes treat both the same way
    model = base.eval()
    g = torch.Generator().manual_seed(2)
    side = 28 if name == "resnet18_small" else 32
    return model, torch.randn(2, 3, side, side, generator=g)


def test_passes_keep_results_and_state_dict_on_cpu():
    for name, units in (("resnet20_cifar", 9), ("resnet18_small", 8)):
        model, x = _student(name)
        with torch.no_grad():
            ref = model(x)
        keys = list(model.state_dict())
        fusion.fuse_eval_bn(model, x)
        assert fusion.fuse_residual_tails(model, x) == units
        assert fusion.fuse_residual_tails(model, x) == 0                 # idempotent
        with torch.no_grad():
            out = model(x)
        assert torch.equal(out, ref) and list(model.state_dict()) == keys
        clone = copy.deepcopy(model)
        unit = next(m for m in clone.modules() if isinstance(m, fusion._FusedUnitMixin))
        own = {id(m) for m in clone.modules()}
        plan = unit._tail_plan
        assert all(id(m) in own for m in [plan.conv, plan.bn1, plan.act] + list(plan.front))


def test_unknown_unit_layouts_are_left_alone():
    class Odd(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1))
            self.activ = torch.nn.ReLU()
            self.resize_identity = False

        def forward(self, x):
            return self.activ(self.body(x))
    m = Odd().eval()
    assert fusion.fuse_residual_tails(m, torch.randn(1, 3, 4, 4)) == 0 and type(m) is Odd


def test_space_to_depth_pass_on_cpu_and_weight_reindexing():
    """The pass only swaps the image-side stride-2 convolution; off the GPU the swapped layer runs its class's own
    forward.  The weight re-indexing itself is checked against the stride-2 convolution with plain torch ops."""
    import torch.nn.functional as F
    model = nets.resnet18_imagenet(num_classes=10).eval()
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = model(x)
    convs_before = [type(m) for m in model.modules() if isinstance(m, torch.nn.Conv2d)]
    assert fusion.space_to_depth_stem(model, x) == 1
    assert fusion.space_to_depth_stem(model, x) == 0
    swapped = [m for m in model.modules() if isinstance(m, fusion._S2DStemMixin)]
    assert len(swapped) == 1 and swapped[0].in_channels == 3 and isinstance(swapped[0], torch.nn.Conv2d)
    assert len([m for m in model.modules() if isinstance(m, torch.nn.Conv2d)]) == len(convs_before)
    with torch.no_grad():
        assert torch.equal(model(x), ref)
    assert fusion.space_to_depth_stem(nets.resnet20_cifar().eval(), torch.randn(1, 3, 32, 32)) == 0   # stride-1 stem

    g = torch.Generator().manual_seed(1)
    for k, p, h in ((7, 3, 16), (3, 1, 10), (5, 2, 12), (4, 1, 8)):
        xx = torch.randn(2, 3, h, h + 2, generator=g)
        w = torch.randn(5, 3, k, k, generator=g)
        n, c, hh, ww = xx.shape
        xp = F.pad(xx, (p, p, p, p))
        xs = xp.view(n, c, (hh + 2 * p) // 2, 2, (ww + 2 * p) // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(
            n, 4 * c, (hh + 2 * p) // 2, (ww + 2 * p) // 2)
        out = F.conv2d(xs, fusion._s2d_weight(w), None, 1, 0)
        want = F.conv2d(xx, w, None, 2, p)
        assert out.shape == want.shape and torch.allclose(out, want, rtol=1e-4, atol=1e-4)


def test_absorbed_modules_do_their_own_work_without_their_batchnorm():
    """An absorbed ReLU / Sequential(ReLU, QuantAct) passes its input through only for the call its BatchNorm has just
    announced.  (a) torchvision's BasicBlock uses ONE ReLU module behind bn1 and behind the residual add: the second
    use must stay a ReLU.  (b) ``convert_sync_batchnorm`` after the pass replaces the fused BatchNorms by fresh
    modules: the absorbed ReLUs must go back to work instead of vanishing from the network."""
    import pytest
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    model = tv.models.resnet18(num_classes=10).eval()
    nets.perturb_bn_stats(model)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        ref = model(x)
    fusion.fuse_eval_bn(model, x)
    assert sum(isinstance(m, fusion.AbsorbedReLU) for m in model.modules()) == 9      # stem + one shared ReLU per block
    with torch.no_grad():
        assert torch.equal(model(x), ref)
    sync = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(model)).eval()
    assert not any(isinstance(m, fusion.FusedEvalBN) for m in sync.modules())
    with torch.no_grad():
        assert torch.allclose(sync(x), ref, rtol=1e-5, atol=1e-6)

    student, x2 = _student("resnet20_cifar")
    with torch.no_grad():
        ref2 = student(x2)
    fusion.fuse_eval_bn(student, x2)
    sync2 = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(student)).eval()
    with torch.no_grad():
        assert torch.equal(student(x2), ref2) and torch.allclose(sync2(x2), ref2, rtol=1e-5, atol=1e-6)


def test_pass_keeps_only_the_safe_part_when_a_batchnorm_output_has_two_readers(monkeypatch):
    """The trace sees who reads a BatchNorm's output first, not who else does: absorbing the ReLU would hand the other
    reader post-ReLU values.  The reader count taken from the autograd graph of the example declines such a pair up
    front; should that analysis be unavailable, the self-check notices, puts the activations back, keeps the BatchNorm
    fusion and warns.  Either way the results are unchanged."""
    import warnings

    import pytest

    class TwoReaders(torch.nn.Module):
        def __init__(self, scale):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 6, 3, padding=1)
            self.bn = torch.nn.BatchNorm2d(6)
            self.relu = torch.nn.ReLU()
            self.mix = torch.nn.Conv2d(6, 4, 1)
            self.scale = scale

        def forward(self, x):
            y = self.bn(self.conv(x))
            return self.mix(self.relu(y) - self.scale * y)        # y is read again after the ReLU

    def build(scale):
        torch.manual_seed(0)
        m = TwoReaders(scale).eval()
        nets.perturb_bn_stats(m)
        x = torch.randn(2, 3, 8, 8)
        with torch.no_grad():
            return m, x, m(x)

    for scale in (0.5, 0.01):                                     # 0.01: far below what an output tolerance could see
        m, x, ref = build(scale)
        with warnings.catch_warnings():
            warnings.simplefilter("error")                        # declined up front: nothing to warn about
            fusion.fuse_eval_bn(m, x)
        assert type(m.bn) is fusion.FusedEvalBN and type(m.relu) is torch.nn.ReLU and m.bn._tail is None
        with torch.no_grad():
            assert torch.equal(m(x), ref)

    monkeypatch.setattr(fusion, "_single_reader", lambda *a: None)   # analysis unavailable: the backstop
    m, x, ref = build(0.5)
    with pytest.warns(UserWarning, match="only the BatchNorms"):
        fusion.fuse_eval_bn(m, x)
    assert type(m.bn) is fusion.FusedEvalBN and type(m.relu) is torch.nn.ReLU and m.bn._tail is None
    with torch.no_grad():
        assert torch.equal(m(x), ref)


def test_units_that_only_look_like_the_layout_are_not_swapped():
    """Same attribute names as the reference BasicBlock, different dataflow (the body is damped by 1 % before the
    add): the per-unit check evaluates the plan's formula with the unit's own modules against what the unit really
    returned, exactly, and declines -- a 1 % deviation is below anything an output tolerance could separate from a
    flipped quantisation code."""
    class Damped(nets.SmallBlock):
        def forward(self, x):
            y = self.relu1(self.bn1(self.conv1(x)))
            y = self.bn2(self.conv2(y)) * 0.99
            return self.relu2(y + self.shortcut(x))
    torch.manual_seed(3)
    model = nets.resnet18_small(3, 9).eval()
    nets.perturb_bn_stats(model)
    blocks = [m for m in model.modules() if type(m) is nets.SmallBlock]
    for b in blocks[::2]:
        b.__class__ = Damped
    x = torch.randn(2, 3, 28, 28)
    with torch.no_grad():
        ref = model(x)
    fusion.fuse_eval_bn(model, x)
    assert fusion.fuse_residual_tails(model, x) == len(blocks) - len(blocks[::2]) == 4
    assert all(not isinstance(b, fusion._FusedUnitMixin) for b in blocks[::2])
    assert all(isinstance(b, fusion._FusedUnitMixin) for b in blocks[1::2])
    with torch.no_grad():
        assert torch.equal(model(x), ref)


def test_global_avgpool_pass_on_cpu_swaps_classes_and_keeps_results():
    """The whole-plane average pool pass is a class swap decided per call: on CPU tensors the module runs its own
    class's forward, the swap is idempotent, survives deepcopy and leaves other pooling modules alone."""
    from torch import nn
    for name in ("resnet20_cifar", "resnet18_small"):
        model, x = _student(name)
        with torch.no_grad():
            ref = model(x)
        pools = [m for m in model.modules() if isinstance(m, (nn.AvgPool2d, nn.AdaptiveAvgPool2d))]
        assert len(pools) == 1
        assert fusion.fuse_global_avgpool(model) == 1 and fusion.fuse_global_avgpool(model) == 0
        assert isinstance(pools[0], (fusion.GlobalAvgPool2d, fusion.GlobalAdaptiveAvgPool2d))
        with torch.no_grad():
            assert torch.equal(copy.deepcopy(model)(x), ref) and torch.equal(model(x), ref)
    other = nn.Sequential(nn.AdaptiveAvgPool2d(3), nn.MaxPool2d(2))
    assert fusion.fuse_global_avgpool(other) == 0 and type(other[0]) is nn.AdaptiveAvgPool2d
