"""north_star (b): ONE pass that updates the running range, emits the fake-quantised tensor and accumulates the
per-channel sum / sum of squares (csrc/fq_calib.cu, ``oodfq_act_calib_stats_forward``).

Reference behaviour being fused: ``QuantAct.forward`` with ``running_stat=True`` (quant_modules.py:80-94) and the two
reductions of the BN-statistics hook on the same tensor (trainer_direct.py:388-393).  The bar: y and the range state
bit-identical to the calibrating forward without statistics (itself bit-identical to the oracle), mean / variance within
1e-5 relative of torch's float64 reductions; both layouts, tensors that fit on chip, tensors whose tail is re-read
through L2, and tensors that take the two-pass fallback.
"""
import numpy as np
import pytest
import torch

from oracle import fq_torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def state():
    return [t.to(DEV) for t in (torch.zeros(1), torch.zeros(1), torch.tensor([0.9]), torch.ones(1))]


def check(x, k, steps=2, onchip=True, rtol=1e-5):
    from ood_dfq_b200 import ops
    sa, sb = state(), state()
    for _ in range(steps):
        y, sums = ops.act_calib_stats_forward(x, k, *sa, onchip=onchip)
        y_ref = ops.act_calib_forward(x, k, *sb, onchip=False)
    assert np.array_equal(torch.cat(sa).cpu().numpy().view(np.int32), torch.cat(sb).cpu().numpy().view(np.int32))
    assert y.stride() == y_ref.stride() and torch.equal(y.view(torch.int32), y_ref.view(torch.int32))
    n, c = x.shape[0], x.shape[1]
    count = x.numel() // c
    mean, var = ops.bn_stats_finalize(sums, None, float(count))
    xd = x.double()
    mean_ref, var_ref = xd.mean([0, 2, 3]), xd.var([0, 2, 3], unbiased=False)
    scale = var_ref.sqrt().max().item() + 1e-30
    assert (mean.double() - mean_ref).abs().max().item() <= rtol * max(mean_ref.abs().max().item(), scale)
    assert ((var.double() - var_ref).abs() <= rtol * var_ref.abs() + 1e-12).all()
    return sums


SHAPES = [(2, 4, 2, 2), (8, 16, 12, 12), (64, 64, 28, 28), (256, 512, 7, 7), (256, 16, 32, 32), (64, 128, 14, 14),
          (64, 512, 4, 4), (3, 8, 5, 7), (256, 256, 14, 14), (256, 128, 28, 28)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("k", [2, 4])
def test_single_pass_range_quant_and_channel_sums(shape, layout, k):
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = torch.relu(torch.randn(shape, generator=g) * 1.3 + 0.2).to(DEV)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    if (x.numel() * 4) % 16 and layout == "nchw":
        pass                                    # numel % 4 != 0 takes the two-pass fallback: same contract
    check(x, k)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_fallback_path_and_forced_two_pass_agree(layout):
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(16, 24, 9, 9, generator=g)).to(DEV)          # C = 24: not a power of two
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    a = check(x, 4)
    b = check(x, 4, onchip=False)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_channel_sums_survive_a_large_offset(layout):
    """|mean| / sigma = 100: the accumulators are taken around a pivot, nothing subtracts two large fp32 numbers."""
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(32, 64, 14, 14, generator=g) * 0.5 + 50.0).to(DEV)
    if layout == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    check(x, 4, rtol=2e-5)


def test_runs_are_deterministic():
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(7)
    for fmt in (torch.contiguous_format, torch.channels_last):
        x = torch.relu(torch.randn(64, 64, 28, 28, generator=g)).to(DEV).contiguous(memory_format=fmt)
        outs = [ops.act_calib_stats_forward(x, 4, *state())[1].clone() for _ in range(3)]
        assert torch.equal(outs[0].view(torch.int64), outs[1].view(torch.int64))
        assert torch.equal(outs[0].view(torch.int64), outs[2].view(torch.int64))


def test_one_launch_when_the_tensor_fits():
    from ood_dfq_b200 import _native, ops
    x = torch.relu(torch.randn(256, 512, 7, 7, device=DEV))
    for fmt in (torch.contiguous_format, torch.channels_last):
        xf = x.contiguous(memory_format=fmt)
        st = state()
        ops.act_calib_stats_forward(xf, 4, *st)                             # cooperative-launch probe, workspace
        _native.reset_launch_count()
        ops.act_calib_stats_forward(xf, 4, *st)
        assert _native.launch_count() == 1


def test_module_collects_statistics_while_calibrating():
    """``QuantAct.collect_channel_stats``: calibrating and frozen forwards both leave the sums of their input; output
    and state follow the oracle module bit for bit."""
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    g = torch.Generator().manual_seed(8)
    ours, ref = qm.QuantAct(4).to(DEV), fq_torch.OracleQuantAct(4)
    ours.collect_channel_stats = True
    for step in range(4):
        if step == 3:
            ours.fix()
            ref.fix()
        x = torch.relu(torch.randn(8, 16, 10, 10, generator=g) * 1.2)
        y, y_ref = ours(x.to(DEV)), ref(x)
        assert np.array_equal(y.cpu().numpy().view(np.int32), y_ref.numpy().view(np.int32))
        for name in ("x_min", "x_max", "beta_t"):
            assert np.array_equal(getattr(ours, name).cpu().numpy().view(np.int32), getattr(ref, name).numpy().view(np.int32))
        mean, var = ours.channel_mean_var()
        np.testing.assert_allclose(mean.cpu().numpy(), x.double().mean([0, 2, 3]).numpy(), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(var.cpu().numpy(), x.double().var([0, 2, 3], unbiased=False).numpy(), rtol=1e-5)
