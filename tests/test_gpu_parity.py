"""GPU parity: the sm_100a kernels (through the C ABI) against the golden vectors and the CPU oracle.

Bars (BASELINE.json north_star): integer codes, dequantised values and range state BIT-EXACT;
BN statistics, loss and gradients within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from conftest import bits
from oracle import bns_torch, fq_torch

pytestmark = pytest.mark.gpu

T = torch.from_numpy
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from ood_dfq_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def qm():
    from ood_dfq_b200.quantization_utils import quant_modules
    return quant_modules


def cu(a):
    t = T(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    return t.to(DEV)


def same_bits(t, ref):
    a = bits(t.detach().cpu().numpy())
    b = bits(ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref)
    return a.shape == b.shape and np.array_equal(a, b)


FROZEN_CASES = [f"{t}_k{k}" for t in ("hw49", "hw16", "flat") for k in (2, 3, 4, 8)] + \
               [f"signed_k{k}" for k in (2, 4, 8)] + ["ties_k4", "degen_k4"]


@pytest.mark.parametrize("case", FROZEN_CASES)
def test_golden_frozen(golden, ops, case):
    g = golden("act_frozen")
    k = int(case.rsplit("_k", 1)[1])
    x, lo, hi = cu(g[f"{case}_x"]), cu(g[f"{case}_lo"]), cu(g[f"{case}_hi"])
    y, codes = ops.fake_quant(x, k, lo, hi, codes=True)
    assert same_bits(y, g[f"{case}_y"])
    assert np.array_equal(codes.cpu().numpy().astype(np.float32), g[f"{case}_codes"])
    y2 = ops.fake_quant(x, k, lo, hi)       # the kernel instance without the codes output
    assert same_bits(y2, g[f"{case}_y"])
    s, z = ops.quant_params(k, lo, hi)
    assert same_bits(s, g[f"{case}_scale"]) and same_bits(z, g[f"{case}_zp"])


def test_golden_helpers(golden, ops):
    """linear_quantize / linear_dequantize / clamp and the param helper, as separate API calls."""
    from ood_dfq_b200.quantization_utils import quant_utils as qu
    g = golden("act_frozen")
    for case in ("hw49_k4", "signed_k8", "ties_k4"):
        k = int(case.rsplit("_k", 1)[1])
        x, lo, hi = cu(g[f"{case}_x"]), cu(g[f"{case}_lo"]), cu(g[f"{case}_hi"])
        s, z = qu.asymmetric_linear_quantization_params(k, lo, hi)
        q = qu.linear_quantize(x, s, z, inplace=False)
        h = 2 ** (k - 1)
        q = qu.clamp(q, -h, h - 1)
        assert same_bits(q, g[f"{case}_codes"])
        y = qu.linear_dequantize(q, s, z, inplace=False)
        assert same_bits(y, g[f"{case}_y"])
        xi = x.clone()
        out = qu.linear_quantize(xi, s, z, inplace=True)
        assert out is xi and same_bits(qu.clamp(xi, -h, h - 1), g[f"{case}_codes"])
        y3 = qu.AsymmetricQuantFunction.apply(x, k, lo, hi)
        assert same_bits(y3, g[f"{case}_y"])
        assert same_bits(qu.find_MSESmallest(x, k, lo, hi), g[f"{case}_y"])


@pytest.mark.parametrize("tag,cls", [("asym", "QuantAct"), ("sym", "QuantAct_DSG")])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_golden_calibrating_sequence(golden, qm, tag, cls, k):
    g = golden("act_calib")
    p = f"{tag}_k{k}_"
    m = getattr(qm, cls)(k).to(DEV)
    for step in range(6):
        if step == 4:
            m.fix()
        if step == 5:
            m.unfix()
        x = cu(g[p + f"x{step}"]).requires_grad_(True)
        y = m(x)
        assert same_bits(y, g[p + f"y{step}"]), step
        state = torch.cat([m.x_min, m.x_max, m.beta_t]).cpu().numpy()
        assert np.array_equal(bits(state), g[p + f"state_bits{step}"]), step
        gout = torch.randn_like(y)
        y.backward(gout)
        assert torch.equal(x.grad, gout)            # identity straight-through estimator


def test_full_precision_passthrough(golden, qm):
    g = golden("act_calib")
    m = qm.QuantAct(4, full_precision_flag=True).to(DEV)
    x = cu(g["fp_x"])
    assert m(x) is x
    state = torch.cat([m.x_min, m.x_max, m.beta_t]).cpu().numpy()
    assert np.array_equal(bits(state), bits(g["fp_state"]))


@pytest.mark.parametrize("tag", ["c3x3", "c1x1", "c7x7", "wide"])
@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("sym", [False, True])
def test_golden_conv_weights(golden, ops, qm, tag, k, sym):
    g = golden("weights")
    p = f"{tag}_k{k}_{'sym' if sym else 'asym'}_"
    w = cu(g[p + "w"])
    (r,) = ops.weight_fq_multi([w], [k], [sym], want_range=True, want_codes=True)
    assert same_bits(r["lo"], g[p + "lo"]) and same_bits(r["hi"], g[p + "hi"])
    assert np.array_equal(r["codes"].cpu().numpy().astype(np.float32), g[p + "codes"])
    assert same_bits(r["wq"], g[p + "wq"])
    # the Function called directly with per-row bounds (generic path)
    from ood_dfq_b200.quantization_utils import quant_utils as qu
    fn = qu.SymmetricQuantFunction_DSG if sym else qu.AsymmetricQuantFunction
    assert same_bits(fn.apply(w, k, cu(g[p + "lo"]), cu(g[p + "hi"])), g[p + "wq"])
    # module forward + STE gradient
    ksz = w.shape[2]
    conv = torch.nn.Conv2d(w.shape[1], w.shape[0], ksz, padding=ksz // 2, bias=(p + "bias") in g)
    with torch.no_grad():
        conv.weight.copy_(T(g[p + "w"]))
        if conv.bias is not None:
            conv.bias.copy_(T(g[p + "bias"]))
    m = (qm.QuantConv2d_DSG if sym else qm.Quant_Conv2d)(k)
    m.set_param(conv)
    m = m.to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    out = m(cu(g[p + "x"]))
    out.square().sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[p + "out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m.weight.grad.cpu().numpy(), g[p + "wgrad"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("sym", [False, True])
def test_golden_linear(golden, qm, k, sym):
    g = golden("weights")
    p = f"lin_k{k}_{'sym' if sym else 'asym'}_"
    lin = torch.nn.Linear(64, 10)
    with torch.no_grad():
        lin.weight.copy_(T(g[p + "w"]))
        lin.bias.copy_(T(g[p + "bias"]))
    m = (qm.QuantLinear_DSG if sym else qm.Quant_Linear)(k)
    m.set_param(lin)
    m = m.to(DEV)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = m(cu(g[p + "x"]))
    out.square().sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), g[p + "out"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m.weight.grad.cpu().numpy(), g[p + "wgrad"], rtol=1e-4, atol=1e-4)


def test_golden_constant_rows(golden, ops):
    g = golden("weights")
    (r,) = ops.weight_fq_multi([cu(g["constrow_w"])], [4], [False])
    assert same_bits(r["wq"], g["constrow_wq"])


def test_golden_mse_range(golden, qm):
    g = golden("act_mse")
    m = qm.QuantAct_MSE(4).to(DEV)
    for step in range(2):
        y = m(cu(g[f"x{step}"]))
        state = torch.cat([m.x_min.reshape(1), m.x_max.reshape(1), m.beta_t.reshape(1)]).cpu().numpy()
        # the search compares L_2.4 scores computed by torch on the GPU; the chosen range and the
        # EMA are fp32 scalar ops -> must agree to rounding of the reduction order
        np.testing.assert_allclose(state, g[f"state{step}"], rtol=1e-6)
        np.testing.assert_allclose(y.cpu().numpy(), g[f"y{step}"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ seeded, oracle-sized
SHAPES = [(64, 64, 56, 56), (32, 512, 7, 7), (16, 3, 33, 31), (5, 7, 3, 1), (1, 1, 1, 1), (1, 1, 1, 9)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("k", [2, 3, 4, 8])
def test_frozen_vs_oracle(ops, shape, k):
    g = torch.Generator().manual_seed(hash((shape, k)) % (2 ** 31))
    x = torch.relu(torch.randn(shape, generator=g) * 1.3)
    lo, hi = torch.zeros(1), (x.max() * 0.9).reshape(1)
    y, codes = ops.fake_quant(cu(x), k, cu(lo), cu(hi), codes=True)
    assert np.array_equal(codes.cpu().numpy().astype(np.float32), fq_torch.codes(x, k, lo, hi).numpy())
    assert same_bits(y, fq_torch.fake_quant(x, k, lo, hi))
    ys = ops.fake_quant(cu(x - 0.7), k, cu(-hi), cu(hi), symmetric=True)
    assert same_bits(ys, fq_torch.fake_quant_symmetric(x - 0.7, k, -hi, hi))


def test_misaligned_views(ops):
    """Storage offsets that break 32-byte (and 16-byte) alignment take the scalar kernel."""
    g = torch.Generator().manual_seed(3)
    base = torch.relu(torch.randn(4099, generator=g))
    lo, hi = torch.zeros(1), torch.ones(1) * 2.0
    for off in (1, 2, 3, 4, 5):
        x = base[off:]
        xg = cu(base)[off:]
        assert same_bits(ops.fake_quant(xg, 4, cu(lo), cu(hi)), fq_torch.fake_quant(x, 4, lo, hi))
        mm = ops.minmax(xg).cpu()
        assert mm[0].item() == x.min().item() and mm[1].item() == x.max().item()


def test_channels_last_and_noncontiguous(ops):
    g = torch.Generator().manual_seed(4)
    x = torch.relu(torch.randn(4, 8, 6, 6, generator=g))
    lo, hi = torch.zeros(1), torch.ones(1) * 1.5
    ref = fq_torch.fake_quant(x, 4, lo, hi)
    y = ops.fake_quant(cu(x).contiguous(memory_format=torch.channels_last), 4, cu(lo), cu(hi))
    assert torch.equal(y.cpu(), ref)
    y = ops.fake_quant(cu(x).transpose(2, 3), 4, cu(lo), cu(hi))
    assert torch.equal(y.cpu(), ref.transpose(2, 3))


def test_nan_semantics(ops, qm):
    """torch.min/max and torch.clamp propagate NaN; so do the kernels."""
    x = torch.relu(torch.randn(3, 4, 5, 5))
    x[1, 2, 3, 4] = float("nan")
    lo, hi = torch.zeros(1), torch.ones(1)
    y = ops.fake_quant(cu(x), 4, cu(lo), cu(hi)).cpu()
    ref = fq_torch.fake_quant(x, 4, lo, hi)
    assert torch.isnan(y[1, 2, 3, 4]) and torch.isnan(ref[1, 2, 3, 4])
    assert torch.equal(torch.nan_to_num(y, 123.0), torch.nan_to_num(ref, 123.0))
    mm = ops.minmax(cu(x)).cpu()
    assert torch.isnan(mm).all() and torch.isnan(x.min()) and torch.isnan(x.max())


def test_cpu_tensor_is_an_error(ops, qm):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qm.QuantAct(4)(torch.randn(2, 3, 4, 4))
    with pytest.raises(RuntimeError):
        ops.fake_quant(torch.randn(4).cuda().double(), 4, torch.zeros(1).cuda(), torch.ones(1).cuda())


def test_empty_tensor(ops):
    y = ops.fake_quant(torch.empty(0, 3, 4, 4, device=DEV), 4, torch.zeros(1, device=DEV), torch.ones(1, device=DEV))
    assert y.shape == (0, 3, 4, 4)
    with pytest.raises(RuntimeError):
        ops.minmax(torch.empty(0, device=DEV))


@pytest.mark.parametrize("sym", [False, True])
def test_weight_bank_multi_tensor(ops, sym):
    """ResNet-18 shaped rows (K = 147 ... 4608) in one launch; every row against the oracle."""
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 3, 7, 7), (64, 64, 3, 3), (128, 64, 1, 1), (128, 128, 3, 3), (256, 256, 3, 3),
              (512, 512, 3, 3), (1000, 512), (10, 64), (9, 2049), (3, 5000)]
    ws = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    res = ops.weight_fq_multi([cu(w) for w in ws], [4] * len(ws), [sym] * len(ws), want_range=True, want_codes=True)
    for w, r in zip(ws, res):
        lo, hi = (fq_torch.row_absmax if sym else fq_torch.row_minmax)(w)
        assert same_bits(r["lo"], lo) and same_bits(r["hi"], hi)
        cfn, ffn = (fq_torch.codes_symmetric, fq_torch.fake_quant_symmetric) if sym else (fq_torch.codes, fq_torch.fake_quant)
        assert np.array_equal(r["codes"].cpu().numpy().astype(np.float32).reshape(w.shape), cfn(w, 4, lo, hi).numpy())
        assert same_bits(r["wq"].view(w.shape), ffn(w, 4, lo, hi))


def test_weight_cache_follows_optimizer(qm):
    conv = torch.nn.Conv2d(8, 16, 3, bias=False)
    m = qm.Quant_Conv2d(4)
    m.set_param(conv)
    m = m.to(DEV)
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    x = torch.randn(2, 8, 6, 6, device=DEV)
    w1 = m.quantized_weight().detach().clone()
    assert m.quantized_weight().data_ptr() == m._wq.data_ptr()       # second call: cache hit
    m(x).square().mean().backward()
    opt.step()
    w2 = m.quantized_weight().detach()
    ref = fq_torch.OracleQuantConv2d(4)
    ref.weight = torch.nn.Parameter(m.weight.detach().cpu())
    assert same_bits(w2, ref.quantized_weight()) and not torch.equal(w1, w2)


# ------------------------------------------------------------------ BN statistics
def test_golden_bns(golden):
    from ood_dfq_b200 import bns
    from test_oracle_golden import tiny_net
    g = golden("bns")
    for flavour in ("trainer", "distill"):
        net = tiny_net(g).to(DEV)
        stat = bns.BNStatLoss(net)
        x = cu(g["x"]).requires_grad_(True)
        torch.backends.cudnn.allow_tf32 = False
        net(x)
        loss = stat.loss()
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"{flavour}_loss"].item(), rtol=1e-5)
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{flavour}_xgrad"], rtol=1e-4, atol=1e-9)
        for i in range(3):
            np.testing.assert_allclose(stat.means()[i].cpu().numpy(), g[f"mean{i}"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(stat.variances()[i].cpu().numpy(), g[f"var{i}"], rtol=1e-5, atol=1e-7)
        lm, lv = stat.parts()
        np.testing.assert_allclose((lm + lv).item(), g["distill_loss"].item(), rtol=1e-5)


@pytest.mark.parametrize("tag", ["off0", "off10", "off100"])
def test_golden_stats_cancellation(golden, tag):
    """|mean|/sigma up to 100: the shifted one-pass variance must still hold 1e-5 (fp64 reference)."""
    from ood_dfq_b200 import bns
    g = golden("bns")
    x = cu(g[f"stat_{tag}_x"])
    mean, var = bns.bn_channel_stats(x)
    np.testing.assert_allclose(mean.cpu().numpy(), g[f"stat_{tag}_mean"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(var.cpu().numpy(), g[f"stat_{tag}_var64"], rtol=1e-5)
    shift = cu(g[f"stat_{tag}_mean"]) + 0.05
    mean, var = bns.bn_channel_stats(x, shift)
    np.testing.assert_allclose(var.cpu().numpy(), g[f"stat_{tag}_var64"], rtol=1e-5)


BN_SHAPES = [(8, 64, 56, 56), (4, 64, 112, 112), (16, 128, 28, 28), (16, 256, 14, 14), (32, 512, 7, 7),
             (32, 512, 4, 4), (3, 5, 7, 9), (2, 3, 1, 1), (5, 130, 7, 7), (2, 6, 70, 70), (4, 2, 33, 35)]


@pytest.mark.parametrize("shape", BN_SHAPES)
def test_channel_stats_and_grad_vs_oracle(shape):
    from ood_dfq_b200 import bns
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g) * 1.5 + 0.3
    c = shape[1]
    rm = torch.randn(c, generator=g) * 0.1
    rv = torch.rand(c, generator=g) + 0.5
    xr = x.clone().requires_grad_(True)
    m_ref, v_ref = bns_torch.channel_stats(xr)
    loss_ref = torch.nn.functional.mse_loss(m_ref, rm) + torch.nn.functional.mse_loss(v_ref, rv)
    loss_ref.backward()
    xg = cu(x).requires_grad_(True)
    mean, var = bns.bn_channel_stats(xg, cu(rm))
    loss = torch.nn.functional.mse_loss(mean, cu(rm)) + torch.nn.functional.mse_loss(var, cu(rv))
    loss.backward()
    np.testing.assert_allclose(mean.detach().cpu().numpy(), m_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var.detach().cpu().numpy(), v_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(loss.item(), loss_ref.item(), rtol=1e-5)
    scale = xr.grad.abs().max().item()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-5 * scale)


def test_fused_stats_with_fake_quant(ops):
    """north_star (b): one read emits the fake-quantised tensor AND the per-channel sums."""
    g = torch.Generator().manual_seed(21)
    for shape in [(8, 64, 28, 28), (16, 512, 7, 7), (4, 16, 70, 70)]:
        x = torch.relu(torch.randn(shape, generator=g))
        lo, hi = torch.zeros(1), (x.max() * 0.8).reshape(1)
        sums, y = ops.bn_stats_forward(cu(x), None, fq=(4, cu(lo), cu(hi)))
        assert same_bits(y, fq_torch.fake_quant(x, 4, lo, hi))
        n = x.numel() // shape[1]
        mean, var = ops.bn_stats_finalize(sums, None, n)
        np.testing.assert_allclose(mean.cpu().numpy(), x.mean([0, 2, 3]).numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(var.cpu().numpy(), x.var([0, 2, 3], unbiased=False).numpy(), rtol=1e-5)


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
@pytest.mark.parametrize("shape", [(256, 64, 112, 112), (256, 512, 7, 7)])
def test_full_size_properties(ops, shape):
    """Size-independent properties at the ImageNet batch-256 shapes (822 MB / 25.7 MB tensors)."""
    torch.manual_seed(0)
    x = torch.relu(torch.randn(shape, device=DEV))
    k = 4
    mm = ops.minmax(x)
    assert mm[0].item() == x.min().item() and mm[1].item() == x.max().item()
    lo, hi = mm[0:1].clone(), (mm[1:2] * 0.7).clone()
    y, codes = ops.fake_quant(x, k, lo, hi, codes=True)
    assert int(codes.min()) >= -8 and int(codes.max()) <= 7
    assert torch.unique(y).numel() <= 2 ** k
    # idempotence: re-quantising the fake-quantised tensor reproduces codes and values
    y2, codes2 = ops.fake_quant(y, k, lo, hi, codes=True)
    assert torch.equal(codes, codes2) and torch.equal(y, y2)
    # monotone: sorting x sorts the codes
    sample = x.flatten()[:: max(1, x.numel() // 1_000_003)]
    order = torch.argsort(sample)
    cs = codes.flatten()[:: max(1, x.numel() // 1_000_003)][order].to(torch.int16)
    assert bool((cs[1:] >= cs[:-1]).all())
    # dequantised values sit on the grid (code + zp)/scale
    s, z = ops.quant_params(k, lo, hi)
    assert torch.equal(y, (codes.float() + z) / s)
    # a slice against the CPU oracle
    sl = x[:2].cpu()
    assert same_bits(y[:2], fq_torch.fake_quant(sl, k, lo.cpu(), hi.cpu()))


def test_quantact_collects_channel_statistics(qm):
    """north_star (b): the QuantAct pass can also deliver per-channel sum / sum of squares of its input."""
    g = torch.Generator().manual_seed(31)
    ref, m = fq_torch.OracleQuantAct(4), qm.QuantAct(4).to(DEV)
    m.collect_channel_stats = True
    for step in range(4):
        if step == 3:
            ref.fix()
            m.fix()
        x = torch.relu(torch.randn(6, 12, 9, 9, generator=g) * (1 + step))
        xg = x.to(DEV).requires_grad_(True)
        y, y_ref = m(xg), ref(x)
        assert same_bits(y, y_ref), step
        assert same_bits(torch.cat([m.x_min, m.x_max, m.beta_t]), torch.cat([ref.x_min, ref.x_max, ref.beta_t]))
        mean, var = m.channel_mean_var()
        np.testing.assert_allclose(mean.cpu().numpy(), x.mean([0, 2, 3]).numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(var.cpu().numpy(), x.var([0, 2, 3], unbiased=False).numpy(), rtol=1e-5)
        gy = torch.randn_like(y)
        y.backward(gy)
        assert torch.equal(xg.grad, gy)
