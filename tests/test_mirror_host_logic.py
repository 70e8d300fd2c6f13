"""Host logic of the mirror modules (ood_dfq_b200/quantization_utils) exercised on CPU.

The modules launch kernels for every tensor operation, so they cannot run here as shipped.  With the launch
wrappers swapped for oracle arithmetic of the same contract (tests/cpu_ops_shim.py, test-only) everything ELSE is
the product's own code: which autograd Function runs in which mode, the in-place updates of ``x_min / x_max /
beta_t``, ``fix`` / ``unfix`` / ``full_precision_flag``, the straight-through gradients, and the ``WeightBank`` that
decides when weights are re-quantised.  Each scenario is run against the LIVE reference classes where the tree is
mounted, against the oracle modules otherwise: bit-identical results are required.
"""
import copy
import os
import sys
import types

import pytest
import torch
from torch import nn

import cpu_ops_shim
from oracle import fq_torch

REF = os.environ.get("OODFQ_REFERENCE", "/root/reference")


def twin_classes():
    """(QuantAct, QuantAct_DSG, QuantAct_MSE, Quant_Conv2d, Quant_Linear, QuantConv2d_DSG, QuantLinear_DSG) of the
    comparison side: the live reference when present."""
    if os.path.isdir(os.path.join(REF, "quantization_utils")):
        import importlib
        pkg = types.ModuleType("_live_reference_qu3")
        pkg.__path__ = [os.path.join(REF, "quantization_utils")]
        sys.modules["_live_reference_qu3"] = pkg
        m = importlib.import_module("_live_reference_qu3.quant_modules")
        return (m.QuantAct, m.QuantAct_DSG, m.QuantAct_MSE, m.Quant_Conv2d, m.Quant_Linear, m.QuantConv2d_DSG,
                m.QuantLinear_DSG), "live reference"
    o = fq_torch
    return (o.OracleQuantAct, o.OracleQuantActSym, o.OracleQuantActMSE, o.OracleQuantConv2d, o.OracleQuantLinear,
            o.OracleQuantConv2dSym, o.OracleQuantLinearSym), "oracle"


TWIN, TWIN_NAME = twin_classes()


@pytest.fixture()
def mirror():
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    with cpu_ops_shim.installed():
        yield qm


def same(a, b):
    return torch.equal(a.detach(), b.detach())


@pytest.mark.parametrize("which", [0, 1, 2])
def test_activation_modules_follow_the_reference_through_every_mode(mirror, which):
    """Calibrating, frozen, full-precision and re-enabled tracking in one sequence; output, gradient and state."""
    ours = [mirror.QuantAct, mirror.QuantAct_DSG, mirror.QuantAct_MSE][which](4)
    ref = TWIN[which](4)
    g = torch.Generator().manual_seed(which)
    script = ["track", "track", "fix", "frozen", "fp", "fp-track", "unfix", "track"]
    for step, what in enumerate(script):
        if what == "fix":
            ours.fix(), ref.fix()
            continue
        if what == "unfix":
            ours.unfix(), ref.unfix()
            ours.full_precision_flag = ref.full_precision_flag = False
            continue
        if what == "fp":
            ours.full_precision_flag = ref.full_precision_flag = True
        if what == "fp-track":
            ours.unfix(), ref.unfix()
        x = (torch.randn(3, 4, 5, 5, generator=g) * (1 + step)).requires_grad_(True)
        xr = x.detach().clone().requires_grad_(True)
        y, yr = ours(x), ref(xr)
        assert same(y, yr), (what, step)
        if what == "fp":
            assert y is x                                     # quant_modules.py:95-96: the input object itself
        cot = torch.randn(y.shape, generator=g)
        (y * cot).sum().backward()
        (yr * cot).sum().backward()
        assert same(x.grad, xr.grad), (what, step)
        for name in ("x_min", "x_max", "beta_t"):
            assert same(getattr(ours, name).reshape(-1), getattr(ref, name).reshape(-1)), (name, what, step)
        assert ours.x_min.shape == (1,) and "x_min" in dict(ours.named_buffers())


@pytest.mark.parametrize("which,layer", [(3, "conv"), (4, "linear"), (5, "conv"), (6, "linear")])
def test_weight_modules_requantise_exactly_when_the_weights_change(mirror, which, layer):
    """Forward + backward + SGD for several steps: the cached fake-quantised weight must follow every optimiser update
    (``_version``), two forwards between updates must reuse it, and results equal the reference's per-forward
    recomputation bit for bit."""
    torch.manual_seed(which)
    src = nn.Conv2d(3, 6, 3, padding=1, bias=True) if layer == "conv" else nn.Linear(12, 5)
    ours = [None, None, None, mirror.Quant_Conv2d, mirror.Quant_Linear, mirror.QuantConv2d_DSG, mirror.QuantLinear_DSG][which](weight_bit=4)
    ref = TWIN[which](weight_bit=4)
    ours.set_param(src), ref.set_param(src)
    opt_o = torch.optim.SGD(ours.parameters(), lr=0.05, momentum=0.9)
    opt_r = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    g = torch.Generator().manual_seed(10)
    calls = []
    real = cpu_ops_shim.weight_fq_multi
    from ood_dfq_b200 import ops
    ops.weight_fq_multi = lambda *a, **k: (calls.append(len(a[0])), real(*a, **k))[1]
    for step in range(4):
        x = torch.randn(2, 3, 6, 6, generator=g) if layer == "conv" else torch.randn(4, 12, generator=g)
        out_a, out_b = ours(x), ours(x * 0.5)                 # two forwards per step, as the QAT iteration does
        ref_a, ref_b = ref(x), ref(x * 0.5)
        assert same(out_a, ref_a) and same(out_b, ref_b), step
        assert len(calls) == step + 1, calls                  # one re-quantisation per optimiser step, not per forward
        opt_o.zero_grad(), opt_r.zero_grad()
        (out_a.square().sum() + out_b.sum()).backward()
        (ref_a.square().sum() + ref_b.sum()).backward()
        assert same(ours.weight.grad, ref.weight.grad) and same(ours.bias.grad, ref.bias.grad)
        opt_o.step(), opt_r.step()
        assert same(ours.weight, ref.weight)
    # a copy of the module owns its own cache; a disabled bank re-quantises on every forward like the reference
    twin = copy.deepcopy(ours)
    with torch.no_grad():
        twin.weight.mul_(2.0)
    assert not same(twin(x), ours(x))
    mirror.WeightBank.enabled = False
    try:
        n = len(calls)
        ours(x), ours(x)
        assert len(calls) == n + 2
    finally:
        mirror.WeightBank.enabled = True
    ours.full_precision_flag = ref.full_precision_flag = True
    assert same(ours(x), ref(x))


def test_quant_utils_functions_match(mirror):
    """The function-level API (quant_utils.py) through the same shim: parameters, quantise, dequantise, STE."""
    from ood_dfq_b200.quantization_utils import quant_utils as qu
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 5, 5, generator=g)
    lo, hi = torch.tensor([-1.5]), torch.tensor([2.25])
    s, z = qu.asymmetric_linear_quantization_params(4, lo, hi)
    s_ref, z_ref = fq_torch.quant_params(4, lo, hi)
    assert same(s, s_ref) and same(z, z_ref)
    q = qu.linear_quantize(x, s, z)
    assert same(q, fq_torch.quantize(x, s_ref, z_ref))
    assert same(qu.linear_dequantize(q, s, z), fq_torch.dequantize(q, s_ref, z_ref))
    keep = x.clone()
    assert qu.linear_quantize(keep, s, z, inplace=True) is keep and same(keep, q)
    xr = x.clone().requires_grad_(True)
    y = qu.AsymmetricQuantFunction.apply(xr, 4, lo, hi)
    assert same(y, fq_torch.fake_quant(x, 4, lo, hi))
    y.sum().backward()
    assert same(xr.grad, torch.ones_like(x))                  # identity STE, no clip mask (quant_utils.py:159-161)
    with pytest.raises(NotImplementedError):
        qu.asymmetric_linear_quantization_params(4, lo, hi, integral_zero_point=False)
    assert same(qu.find_MSESmallest(x, 4, lo, hi), fq_torch.fake_quant(x, 4, lo, hi))
    assert abs(qu.lp_loss(x, x * 0.5, p=2.4, reduction="all").item() - (x * 0.5).abs().pow(2.4).mean().item()) < 1e-6


@pytest.mark.parametrize("net_name,side,expect", [
    # (BN-fed tails absorbed, tails behind a residual add, stem pools absorbed, residual units)
    ("resnet18_imagenet", 224, (9, 8, 1, 8)),
    ("resnet20_cifar", 32, (10, 9, 0, 9)),
    ("resnet18_small", 28, (9, 8, 0, 8)),
])
def test_fusion_passes_on_a_quantised_student(mirror, net_name, side, expect):
    """The passes trace the QUANTISED student (Sequential(ReLU(inplace), QuantAct) tails) -- possible on CPU with the
    shim: every BatchNorm that directly feeds a tail absorbs it, the tails behind residual adds become the fused
    ReLU+QuantAct pair, the ImageNet stem absorbs its max-pool, every residual unit is recognised; off the GPU the
    fused modules run the original chain, so outputs, ranges and keys are unchanged."""
    from ood_dfq_b200 import fusion, nets, surgery
    torch.manual_seed(2)
    base = nets.resnet18_small(3, 9) if net_name == "resnet18_small" else getattr(nets, net_name)(num_classes=10)
    nets.perturb_bn_stats(base)
    student = surgery.quantize_model(base, 4, 4).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, side, side, generator=g)
    with torch.no_grad():
        student(x)                                            # one calibrating forward, then freeze
    surgery.freeze_model(student)
    with torch.no_grad():
        ref = student(x)
    keys = list(student.state_dict())
    state = {k: v.clone() for k, v in student.state_dict().items()}
    fusion.fuse_eval_bn(student, x)
    units = fusion.fuse_residual_tails(student, x)
    fusion.space_to_depth_stem(student, x)
    count = lambda cls: sum(type(m) is cls for m in student.modules())
    assert (count(fusion.AbsorbedTail), count(fusion.FusedReLUQuant), count(fusion.AbsorbedPool), units) == expect
    with torch.no_grad():
        out = student(x)
    assert torch.equal(out, ref) and list(student.state_dict()) == keys
    assert all(torch.equal(v, state[k]) for k, v in student.state_dict().items())
    # calibration keeps working through the fused modules (they step aside while a QuantAct tracks its range)
    surgery.unfreeze_model(student)
    plain = surgery.quantize_model(base, 4, 4).eval()
    plain.load_state_dict(state)
    with torch.no_grad():
        assert torch.equal(student(x * 1.5), plain(x * 1.5))
    assert all(torch.equal(a, b) for a, b in zip(student.state_dict().values(), plain.state_dict().values()))


# ------------------------------------------------------------------------------------------------ BN-statistics loss
def _bn_net(seed=4):
    torch.manual_seed(seed)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                        nn.Conv2d(8, 12, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(12), nn.ReLU(),
                        nn.Conv2d(12, 5, 1, bias=False), nn.BatchNorm2d(5)).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for m in net:
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return net


def test_bn_stat_loss_manager_matches_the_hook_and_loss_of_the_reference(mirror):
    """``bns.BNStatLoss`` (hooks, packed sums, token plumbing, fused backward) with shimmed kernels against the
    reference's hook + loss lines restated by the oracle (trainer flavour and distillation parts)."""
    from ood_dfq_b200 import bns
    from oracle import bns_torch
    net = _bn_net()
    x = torch.randn(6, 3, 10, 10, generator=torch.Generator().manual_seed(9)) * 1.5 + 0.3
    xr = x.clone().requires_grad_(True)
    tap = bns_torch.StatTap(net)
    out_ref = net(xr)
    loss_ref = tap.loss("trainer")
    (out_ref.square().mean() + 0.1 * loss_ref).backward()
    tap.remove()
    xo = x.clone().requires_grad_(True)
    mgr = bns.BNStatLoss(net)
    out = net(xo)
    loss = mgr.loss()
    (out.square().mean() + 0.1 * loss).backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item())
    assert torch.allclose(xo.grad, xr.grad, rtol=1e-5, atol=1e-8)
    m_term, v_term = mgr.parts()
    assert abs((m_term + v_term).item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item())
    for a, b in zip(mgr.means(), tap.means):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)
    for a, b in zip(mgr.variances(), tap.vars):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)
    # a pass in which a hooked module runs twice is refused, and the manager recovers
    net(x), net[1](torch.randn(2, 8, 4, 4))
    with pytest.raises(RuntimeError, match="fired"):
        mgr.loss()
    net(x)
    assert abs(mgr.loss().item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item())
    mgr.remove()
    # the function-level drop-in for the two reductions of the hook
    xs = x[:, :3].clone().requires_grad_(True)
    mean, var = bns.bn_channel_stats(xs)
    (mean.sum() + 2 * var.sum()).backward()
    xt = x[:, :3].clone().requires_grad_(True)
    mt, vt = bns_torch.channel_stats(xt)
    (mt.sum() + 2 * vt.sum()).backward()
    assert torch.allclose(mean, mt, rtol=1e-5, atol=1e-6) and torch.allclose(var, vt, rtol=1e-5)
    assert torch.allclose(xs.grad, xt.grad, rtol=1e-5, atol=1e-8)


def test_quantact_with_channel_statistics_in_the_same_pass(mirror):
    """north_star (b): ``collect_channel_stats`` makes the quantising pass also leave per-channel sums of its input.
    Output, range state and gradient must be those of the plain module in both modes; the statistics those of the hook."""
    ours, plain = mirror.QuantAct(4), TWIN[0](4)
    ours.collect_channel_stats = True
    g = torch.Generator().manual_seed(8)
    for step in range(4):
        if step == 2:
            ours.fix(), plain.fix()
        x = torch.relu(torch.randn(5, 6, 7, 7, generator=g) * (1 + step)).requires_grad_(True)
        xr = x.detach().clone().requires_grad_(True)
        y, yr = ours(x), plain(xr)
        assert same(y, yr), step
        for name in ("x_min", "x_max", "beta_t"):
            assert same(getattr(ours, name).reshape(-1), getattr(plain, name).reshape(-1)), (name, step)
        y.sum().backward(), yr.sum().backward()
        assert same(x.grad, xr.grad)
        mean, var = ours.channel_mean_var()
        assert torch.allclose(mean, x.detach().mean([0, 2, 3]), rtol=1e-5, atol=1e-6)
        assert torch.allclose(var, x.detach().var([0, 2, 3], unbiased=False), rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError, match="no statistics"):
        mirror.QuantAct(4).channel_mean_var()


def test_bn_stat_loss_sees_running_statistics_rewritten_through_data(mirror):
    """The reference's BN-statistic delta correction rewrites running statistics with ``running_mean.data.copy_(...)``
    (trainer_direct.py:292-297), which does not bump the buffer's version counter; the loss must follow anyway."""
    from ood_dfq_b200 import bns
    from oracle import bns_torch
    net = _bn_net()
    x = torch.randn(4, 3, 10, 10, generator=torch.Generator().manual_seed(2))
    mgr = bns.BNStatLoss(net)
    net(x)
    before = mgr.loss().item()
    for m in net:
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.data.copy_(m.running_mean * 0.5 + 1.0)
            m.running_var.data.copy_(m.running_var * 2.0)
    net(x)
    after = mgr.loss().item()
    mgr.remove()
    tap = bns_torch.StatTap(net)
    net(x)
    want = tap.loss("trainer").item()
    assert abs(after - want) <= 1e-6 * abs(want) and abs(after - before) > 1e-3


class _PretendCudaNHWC:
    """What ``_fused_setup`` asks of its input, answered like a channels_last fp32 CUDA tensor (decision logic only)."""
    is_cuda, dtype = True, torch.float32

    def dim(self):
        return 4

    def is_contiguous(self, memory_format=torch.contiguous_format):
        return memory_format is torch.channels_last


def test_fused_units_step_aside_for_foreign_hooks_and_calibration(mirror):
    """The decision whether a residual unit takes its one-kernel tail is host logic: it must say yes for a frozen,
    eval-mode, hook-free unit, and no as soon as (a) a BatchNorm it would bypass carries a forward hook -- the
    reference's BN-statistic delta correction hangs one on EVERY BatchNorm (trainer_direct.py:176-199, :215-240) --,
    (b) its QuantAct tracks the range again, (c) the model trains; the trainer's own feature hook on ``body`` is the
    one hook it serves itself."""
    from ood_dfq_b200 import fusion, nets, step, surgery
    torch.manual_seed(2)
    base = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(base)
    student = surgery.quantize_model(base, 4, 4).eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        student(x)
    surgery.freeze_model(student)
    fusion.fuse_eval_bn(student, x)
    assert fusion.fuse_residual_tails(student, x) == 9
    units = [m for m in student.modules() if isinstance(m, fusion._FusedUnitMixin)]
    fake = _PretendCudaNHWC()
    assert all(u._fused_setup(fake) is not None for u in units)
    assert all(u._fused_setup(torch.zeros(1, 16, 4, 4)) is None for u in units)          # a CPU tensor: never

    u = units[3]
    plan = u._tail_plan
    h = plan.bn1.register_forward_hook(lambda m, i, o: None)                              # (a) BSDC-style statistics hook
    assert u._fused_setup(fake) is None and units[4]._fused_setup(fake) is not None
    h.remove()
    assert u._fused_setup(fake) is not None
    qact = plan.act[1]
    qact.unfix()                                                                          # (b) calibrating again
    assert u._fused_setup(fake) is None
    qact.fix()
    student.train()                                                                       # (c) training mode
    assert u._fused_setup(fake) is None
    student.eval()
    tap = step.FeatureTap(student, (nets.ResUnit,), fused=True)                           # the trainer's feature hook
    setup = u._fused_setup(fake)
    assert setup is not None and setup[2] == [tap]
    other = u.body.register_forward_hook(lambda m, i, o: None)                            # any other hook there: no
    assert u._fused_setup(fake) is None
    other.remove()
    qact.full_precision_flag = True                                                       # fp activations: fused, no quantiser
    assert u._fused_setup(fake)[1] is None


def test_fused_units_notice_when_their_modules_were_rebuilt(mirror):
    """``convert_sync_batchnorm`` after the passes replaces every BatchNorm object; a fused unit whose plan still points
    at the old ones must take its class's own forward (which goes through the attributes)."""
    from ood_dfq_b200 import fusion, nets, surgery
    torch.manual_seed(2)
    student = surgery.quantize_model(nets.resnet20_cifar(num_classes=10), 4, 4).eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        student(x)
    surgery.freeze_model(student)
    fusion.fuse_eval_bn(student, x)
    fusion.fuse_residual_tails(student, x)
    with torch.no_grad():
        ref = student(x)
    fake = _PretendCudaNHWC()
    units = [m for m in student.modules() if isinstance(m, fusion._FusedUnitMixin)]
    assert all(u._fused_setup(fake) is not None for u in units)
    sync = torch.nn.SyncBatchNorm.convert_sync_batchnorm(student).eval()
    units = [m for m in sync.modules() if isinstance(m, fusion._FusedUnitMixin)]
    assert len(units) == 9 and all(u._fused_setup(fake) is None for u in units)
    with torch.no_grad():
        assert torch.allclose(sync(x), ref, rtol=1e-5, atol=1e-6)


def _random_net(rng, width=8):
    """A small CNN assembled from the patterns the passes have to cope with (and some they must leave alone)."""
    class Res(nn.Module):
        def __init__(self, c, style):
            super().__init__()
            self.conv1, self.bn1 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c)
            self.conv2, self.bn2 = nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c)
            self.style = style
            if style == "shared":                              # torchvision BasicBlock: one ReLU module, two sites
                self.relu = nn.ReLU(inplace=True)
            else:
                self.relu1, self.relu2 = nn.ReLU(inplace=True), nn.ReLU(inplace=bool(rng.integers(2)))

        def forward(self, x):
            r1 = self.relu if self.style == "shared" else self.relu1
            r2 = self.relu if self.style == "shared" else self.relu2
            out = self.bn2(self.conv2(r1(self.bn1(self.conv1(x)))))
            if self.style == "inplace":
                out += x                                       # reference models.py:40-41
            else:
                out = out + x
            return r2(out)

    class TwoReaders(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv, self.bn, self.relu = nn.Conv2d(c, c, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU()

        def forward(self, x):
            y = self.bn(self.conv(x))
            return self.relu(y) + 0.25 * y

    layers, c = [nn.Conv2d(3, width, 3, padding=1, bias=False), nn.BatchNorm2d(width), nn.ReLU(inplace=True)], width
    for _ in range(int(rng.integers(2, 6))):
        kind = rng.choice(["cbr", "cb", "cbr6", "pool", "res_out", "res_inplace", "res_shared", "two"])
        if kind == "cbr":
            layers += [nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=bool(rng.integers(2)))]
        elif kind == "cb":
            layers += [nn.Conv2d(c, c, 1, bias=False), nn.BatchNorm2d(c)]
        elif kind == "cbr6":
            layers += [nn.Conv2d(c, c, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU6()]
        elif kind == "pool":
            layers += [nn.MaxPool2d(3, 2, 1)]
        elif kind == "two":
            layers += [TwoReaders(c)]
        else:
            layers += [Res(c, kind.split("_")[1])]
    layers += [nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(c, 5)]
    return nn.Sequential(*layers)


@pytest.mark.parametrize("seed", range(24))
def test_fusion_passes_never_change_a_random_network(mirror, seed):
    """Full-precision and quantised: whatever the layout, after ``fuse_eval_bn`` + ``fuse_residual_tails`` the network
    computes what it computed before (the passes may decline, warn, or keep only the BatchNorm part)."""
    import warnings

    import numpy as np

    from ood_dfq_b200 import fusion, nets, surgery
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    net = _random_net(rng).eval()
    nets.perturb_bn_stats(net, seed=seed)
    x = torch.randn(3, 3, 16, 16)
    for quantised in (False, True):
        model = copy.deepcopy(net)
        if quantised:
            model = surgery.quantize_model(model, 4, 4).eval()
            with torch.no_grad():
                model(x)
            surgery.freeze_model(model)
        with torch.no_grad():
            ref = model(x)
        keys = list(model.state_dict())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fusion.fuse_eval_bn(model, x)
            fusion.fuse_residual_tails(model, x)
        with torch.no_grad():
            out = model(x)
        assert torch.equal(out, ref), (seed, quantised)
        assert list(model.state_dict()) == keys
        twin = copy.deepcopy(model)
        with torch.no_grad():
            assert torch.equal(twin(x), ref)


def test_passes_do_not_move_the_ranges_of_a_calibrating_student(mirror):
    """Applied before ``freeze_model`` the passes still run the example through the model several times; every
    calibrating QuantAct must come out with exactly the state it went in with."""
    from ood_dfq_b200 import fusion, nets, surgery
    torch.manual_seed(4)
    student = surgery.quantize_model(nets.resnet20_cifar(num_classes=10), 4, 4).eval()
    x = torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        student(x)
    before = {k: v.clone() for k, v in student.state_dict().items()}
    fusion.fuse_eval_bn(student, x)
    assert fusion.fuse_residual_tails(student, x) == 9
    after = student.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)
    assert any(k.endswith("beta_t") and 0 < float(v) < 1 for k, v in after.items())      # it WAS calibrating


def test_qat_step_on_a_fused_mirror_student_equals_the_oracle_step(mirror):
    """``step.QATStep`` (flat gradient buffer, pruned backward, feature taps) over a mirror-built student that went
    through every fusion pass, against the same step over the oracle-built student: identical losses and weights for
    several iterations (off the GPU the fused modules run the chains they replace; the bookkeeping around them --
    class swaps, absorbed containers, WeightBank -- is what is exercised)."""
    from ood_dfq_b200 import fusion, nets, step, surgery
    torch.manual_seed(6)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    g = torch.Generator().manual_seed(7)
    calib = torch.randn(4, 3, 32, 32, generator=g)
    batches = [torch.randn(4, 3, 32, 32, generator=g) for _ in range(3)]
    results = []
    for namespace, fuse in ((fq_torch, False), (None, True)):
        student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=namespace).eval()
        t = copy.deepcopy(teacher).eval()
        with torch.no_grad():
            student(calib)
        surgery.freeze_model(student, namespace)
        if fuse:
            for net in (student, t):
                fusion.fuse_eval_bn(net, calib[:2])
                assert fusion.fuse_residual_tails(net, calib[:2]) == 9
        qat = step.QATStep(student, t, lr=1e-4, unit_types=(nets.ResUnit,))
        losses = [qat(b).clone() for b in batches]
        results.append((losses, [p.detach().clone() for p in student.parameters()]))
    (l_a, p_a), (l_b, p_b) = results
    assert all(torch.equal(a, b) for a, b in zip(l_a, l_b))
    assert all(torch.equal(a, b) for a, b in zip(p_a, p_b))


def test_weight_cache_is_dropped_by_state_dict_loads_data_writes_through_apply_and_the_step(mirror):
    """The fake-quantised weight is cached per parameter version.  Writes that bypass the version counter must still
    be seen where the package can see them: ``load_state_dict`` (copies under no_grad -- bumps), ``Module._apply``
    (``.to()`` / ``.double().float()`` rewrite ``.data``: no bump) and the optimiser update inside ``QATStep.apply``
    (torch's fused SGD does not bump); the reference re-quantises on every forward and would notice all of them."""
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(4, 6, 3, bias=False)
    q, ref = mirror.Quant_Conv2d(4), TWIN[2](4) if len(TWIN) > 2 else None
    q.set_param(conv)
    x = torch.randn(2, 4, 8, 8)
    y0 = q(x)
    other = {"weight": torch.randn_like(conv.weight)}
    q.load_state_dict(other)
    y1 = q(x)
    want = torch.nn.functional.conv2d(x, cpu_ops_shim.weight_fq_multi([other["weight"]], [4], [False])[0]["wq"])
    assert same(y1, want) and not same(y1, y0)
    # a write through .data that nothing can observe keeps serving the cache (documented) ...
    q.weight.data.mul_(2.0)
    assert same(q(x), y1)
    # ... until something the package sees happens: Module._apply
    q._apply(lambda t: t)
    want2 = torch.nn.functional.conv2d(x, cpu_ops_shim.weight_fq_multi([q.weight.detach()], [4], [False])[0]["wq"])
    assert same(q(x), want2) and not same(want2, y1)
