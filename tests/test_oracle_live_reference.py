"""Randomised cross-check of the three oracle restatements against the LIVE reference.

The committed golden vectors (tests/golden, tools/make_golden.py) pin the oracle at fixed seeds.  In the build
container the reference tree itself is importable, so here hypothesis draws shapes, bit-widths, ranges and data
(special values included) and every restatement must reproduce ``quantization_utils`` bit for bit.  The GPU box
has no ``/root/reference``: the module skips there (nothing under ``-m gpu`` reads the reference).
"""
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("OODFQ_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the operator package: the mounted reference tree, else the unmodified copy oracle/make_ref.py stages under oracle/_ref
# (git-ignored; it travels to the GPU box, so these cross-checks run there too)
QU_DIR = os.path.join(REF, "quantization_utils")
if not os.path.isdir(QU_DIR):
    QU_DIR = os.path.join(ROOT, "oracle", "_ref", "quantization_utils")
if not os.path.isfile(os.path.join(QU_DIR, "quant_modules.py")):
    pytest.skip("neither the reference tree nor its staged copy (oracle/_ref) is present: golden vectors cover the pin",
                allow_module_level=True)
HAVE_TREE = os.path.isdir(os.path.join(REF, "data_generate"))

hyp = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from conftest import bits  # noqa: E402
from oracle import fq_c, fq_numpy, fq_torch  # noqa: E402


def _reference():
    """The reference package under a private name, so that the product's drop-in ``quantization_utils`` (which other
    tests may have installed as a top-level module) is never what gets imported here."""
    import importlib
    import types
    pkg = types.ModuleType("_live_reference_qu")
    pkg.__path__ = [QU_DIR]                                       # quant_modules.py:28 imports `.quant_utils`
    sys.modules["_live_reference_qu"] = pkg
    return (importlib.import_module("_live_reference_qu.quant_utils"),
            importlib.import_module("_live_reference_qu.quant_modules"))


RU, RM = _reference()
torch.set_num_threads(1)

COMMON = dict(deadline=None, max_examples=200, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
SPECIALS = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, -1e-30, 3.4e38, -3.4e38, 0.5, 1.5, 2.5, -0.5, -1.5],
                    dtype=np.float32)


def draw_data(seed, shape, spread, shift, specials):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(shape) * spread + shift).astype(np.float32)
    if specials:
        flat = x.reshape(-1)
        idx = rng.integers(0, flat.size, size=min(flat.size, 6))
        flat[idx] = rng.choice(SPECIALS, size=len(idx))
    return x


shapes4 = st.tuples(st.integers(1, 4), st.integers(1, 9), st.integers(1, 9), st.integers(1, 9))
ranges = st.one_of(
    st.tuples(st.floats(-8, 0, width=32), st.floats(0, 8, width=32)),
    st.tuples(st.just(0.0), st.floats(2.0 ** -20, 50, width=32)),                 # post-ReLU: data-min = 0
    st.tuples(st.floats(-3, 3, width=32), st.floats(-3, 3, width=32)),      # inverted / degenerate spans too
    st.tuples(st.floats(-2.0 ** -30, 2.0 ** -30, width=32), st.floats(-2.0 ** -30, 2.0 ** -30, width=32)),
)


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), shape=shapes4, k=st.sampled_from([1, 2, 3, 4, 5, 6, 8, 12, 16]), rng=ranges,
       spread=st.sampled_from([1e-3, 1.0, 30.0]), specials=st.booleans())
def test_frozen_fake_quant_matches_the_live_reference(seed, shape, k, rng, spread, specials):
    """AsymmetricQuantFunction.forward (quant_utils.py:138-157) with a scalar range: codes and values, three oracles."""
    lo, hi = np.float32(rng[0]), np.float32(rng[1])
    x = draw_data(seed, shape, spread, 0.0, specials)
    xt, lot, hit = torch.from_numpy(x), torch.tensor([lo]), torch.tensor([hi])
    s, z = RU.asymmetric_linear_quantization_params(k, lot, hit)
    h = 2 ** (k - 1)
    q_ref = torch.clamp(RU.linear_quantize(xt, s, z, inplace=False), -h, h - 1).numpy()
    y_ref = RU.AsymmetricQuantFunction.apply(xt, k, lot, hit).numpy()

    s_o, z_o = fq_torch.quant_params(k, lot, hit)
    assert np.array_equal(bits(s_o.numpy()), bits(s.numpy())) and np.array_equal(bits(z_o.numpy()), bits(z.numpy()))
    assert np.array_equal(bits(fq_torch.codes(xt, k, lot, hit).numpy()), bits(q_ref))
    assert np.array_equal(bits(fq_torch.fake_quant(xt, k, lot, hit).numpy()), bits(y_ref))

    s_n, z_n = fq_numpy.quant_params(k, lo, hi)
    assert np.array_equal(bits(s_n), bits(s.numpy().reshape(()))) and np.array_equal(bits(z_n), bits(z.numpy().reshape(())))
    nan = np.isnan(y_ref)
    with np.errstate(all="ignore"):
        q_n, y_n = fq_numpy.codes(x, k, lo, hi), fq_numpy.fake_quant(x, k, lo, hi)
        y_c, q_c = fq_c.fake_quant(x, k, lo, hi)
    for q, y in ((q_n, y_n), (q_c, y_c)):           # NaN payloads are not part of the contract, NaN-ness is
        assert np.array_equal(np.isnan(y), nan) and np.array_equal(np.isnan(q), np.isnan(q_ref))
        assert np.array_equal(bits(y)[~nan], bits(y_ref)[~nan])
        # codes are integers: the sign of a zero code is unobservable (k = 1 clamps at an upper bound of 0, where
        # torch.clamp, np.minimum and fminf each keep a different zero); the dequantised value above IS compared by bits
        assert np.array_equal(q[~np.isnan(q_ref)], q_ref[~np.isnan(q_ref)])


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), rows=st.integers(1, 12), cols=st.tuples(st.integers(1, 6), st.integers(1, 3), st.integers(1, 3)),
       k=st.sampled_from([2, 3, 4, 8]), sym=st.booleans(), linear=st.booleans())
def test_weight_modules_match_the_live_reference(seed, rows, cols, k, sym, linear):
    """Quant_Conv2d / Quant_Linear and their DSG twins (quant_modules.py:188-281, :389-481): per-row ranges,
    fake-quantised weight and the STE gradient reaching the Parameter."""
    torch.manual_seed(seed % (2 ** 31))
    if linear:
        layer = torch.nn.Linear(cols[0] * cols[1], rows, bias=True)
        ref = (RM.QuantLinear_DSG if sym else RM.Quant_Linear)(weight_bit=k)
        ours = (fq_torch.OracleQuantLinearSym if sym else fq_torch.OracleQuantLinear)(k)
        x = torch.randn(3, cols[0] * cols[1])
    else:
        layer = torch.nn.Conv2d(cols[0], rows, (cols[1], cols[2]), padding=1, bias=False)
        ref = (RM.QuantConv2d_DSG if sym else RM.Quant_Conv2d)(weight_bit=k)
        ours = (fq_torch.OracleQuantConv2dSym if sym else fq_torch.OracleQuantConv2d)(k)
        x = torch.randn(2, cols[0], 5, 5)
    ref.set_param(layer)
    ours.set_param(layer)
    out_ref, out = ref(x), ours(x)
    assert np.array_equal(bits(out.detach().numpy()), bits(out_ref.detach().numpy()))
    out_ref.square().sum().backward()
    out.square().sum().backward()
    assert np.array_equal(bits(ours.weight.grad.numpy()), bits(ref.weight.grad.numpy()))
    # the C and numpy restatements on the same rows
    w = layer.weight.detach().numpy().reshape(rows, -1)
    lo, hi = fq_c.row_ranges(w, symmetric=sym)
    wq_ref = ours.quantized_weight().detach().numpy().reshape(rows, -1)
    y_c, _ = fq_c.fake_quant(w, k, lo, hi, symmetric=sym)
    assert np.array_equal(bits(y_c), bits(wq_ref))
    assert np.array_equal(bits(fq_numpy.fake_quant(w, k, lo, hi, symmetric=sym)), bits(wq_ref))


@settings(**{**COMMON, "max_examples": 100})
@given(seed=st.integers(0, 2 ** 31), shape=shapes4, k=st.sampled_from([2, 4, 8]), sym=st.booleans(),
       steps=st.integers(1, 7), shift=st.sampled_from([0.0, 2.0, -5.0]), fix_at=st.integers(0, 7))
def test_calibrating_sequences_match_the_live_reference(seed, shape, k, sym, steps, shift, fix_at):
    """QuantAct / QuantAct_DSG over several forwards incl. fix(): state (x_min, x_max, beta_t) and output, bit-exact;
    the numpy and C restatements follow the same state."""
    ref = (RM.QuantAct_DSG if sym else RM.QuantAct)(activation_bit=k)
    ours = (fq_torch.OracleQuantActSym if sym else fq_torch.OracleQuantAct)(k)
    st_np = (np.float32(0), np.float32(0), np.float32(1))
    st_c = np.array([0, 0, 1], dtype=np.float32)
    for step in range(steps):
        if step == fix_at:
            ref.fix()
            ours.fix()
        x = draw_data(seed + step, shape, 1.0 + step, shift, False)
        xt = torch.from_numpy(x)
        y_ref, y = ref(xt), ours(xt)
        assert np.array_equal(bits(y.numpy()), bits(y_ref.numpy()))
        for name in ("x_min", "x_max", "beta_t"):
            assert np.array_equal(bits(getattr(ours, name).numpy()), bits(getattr(ref, name).numpy())), (name, step)
        if step < fix_at:
            st_np = fq_numpy.range_update(*st_np[:2], np.float32(0.9), st_np[2], x.min(), x.max(), symmetric=sym)
            st_c = fq_c.range_update(st_c, 0.9, x.min(), x.max(), symmetric=sym)
            want = np.array([ref.x_min.item(), ref.x_max.item(), ref.beta_t.item()], dtype=np.float32)
            assert np.array_equal(bits(np.array(st_np, dtype=np.float32)), bits(want))
            assert np.array_equal(bits(st_c), bits(want))


@settings(**{**COMMON, "max_examples": 12})
@given(seed=st.integers(0, 2 ** 31), shape=shapes4, k=st.sampled_from([2, 4, 8]), steps=st.integers(1, 3),
       relu=st.booleans())
def test_mse_searched_ranges_match_the_live_reference(seed, shape, k, steps, relu):
    """QuantAct_MSE (quant_modules.py:98-186): 80-candidate clip search + plain EMA, state and output bit-exact."""
    ref, ours = RM.QuantAct_MSE(activation_bit=k), fq_torch.OracleQuantActMSE(k)
    for step in range(steps):
        x = draw_data(seed + step, shape, 1.0 + step, 0.3, False)
        xt = torch.from_numpy(np.maximum(x, 0) if relu else x)
        y_ref, y = ref(xt), ours(xt)
        assert np.array_equal(bits(y.numpy()), bits(y_ref.numpy()))
        for name in ("x_min", "x_max", "beta_t"):
            assert np.array_equal(bits(getattr(ours, name).numpy()), bits(getattr(ref, name).numpy())), (name, step)


@settings(**{**COMMON, "max_examples": 25})
@given(seed=st.integers(0, 2 ** 31), shape=st.tuples(st.integers(2, 5), st.integers(1, 6), st.integers(1, 7), st.integers(1, 7)),
       offset=st.sampled_from([0.0, 3.0, -40.0]), flavour=st.sampled_from(["trainer", "distill"]))
def test_bn_statistics_hook_and_loss_match_the_live_reference(seed, shape, offset, flavour):
    """The reference's BN hook (data_generate/distill_data.py:69-78, same body as trainer_direct.py:388-397) and
    the loss lines around it (:252-265 / trainer_direct.py:473-486) against the oracle's StatTap: statistics, loss
    and the gradient reaching the input, bit for bit (both are the same ATen calls in the same order)."""
    if not HAVE_TREE:
        pytest.skip("needs data_generate/distill_data.py of the mounted reference tree")
    sys.path.insert(0, REF)
    try:
        from data_generate.distill_data import DistillData
    finally:
        sys.path.remove(REF)
    from oracle import bns_torch
    torch.manual_seed(seed % (2 ** 31))
    c = shape[1]
    net = torch.nn.Sequential(torch.nn.BatchNorm2d(c), torch.nn.Conv2d(c, c + 1, 1), torch.nn.BatchNorm2d(c + 1)).eval()
    for m in net:
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features) * 0.3 + offset)
            m.running_var.copy_(torch.rand(m.num_features) + 0.5)
    x = torch.from_numpy(draw_data(seed, shape, 1.3, offset, False))
    dd = DistillData()
    handles = [m.register_forward_hook(dd.hook_fn_forward) for m in net if isinstance(m, torch.nn.BatchNorm2d)]
    xr = x.clone().requires_grad_(True)
    net(xr)
    mse, n_l = torch.nn.MSELoss(), len(dd.mean_list)
    if flavour == "trainer":             # the calls of trainer_direct.py:474-484 (the file itself has a TabError)
        loss_ref = torch.zeros(1)
        for i in range(n_l):
            loss_ref += mse(dd.mean_list[i], dd.teacher_running_mean[i]) + mse(dd.var_list[i], dd.teacher_running_var[i])
        loss_ref = loss_ref / n_l
    else:                                # distill_data.py:252-265
        ml, vl = torch.zeros(1), torch.zeros(1)
        for i in range(n_l):
            ml += mse(dd.mean_list[i], dd.teacher_running_mean[i].detach())
            vl += mse(dd.var_list[i], dd.teacher_running_var[i].detach())
        loss_ref = ml / n_l + vl / n_l
    loss_ref.backward()
    for h in handles:
        h.remove()
    xo = x.clone().requires_grad_(True)
    tap = bns_torch.StatTap(net)
    net(xo)
    loss = tap.loss(flavour)
    loss.backward()
    tap.remove()
    for i in range(n_l):
        assert torch.equal(tap.means[i], dd.mean_list[i]) and torch.equal(tap.vars[i], dd.var_list[i])
    assert torch.equal(loss, loss_ref) and torch.equal(xo.grad, xr.grad)
    # the closed form the GPU backward implements, first layer only (its input IS x)
    g = bns_torch.bns_input_grad(x.double(), net[0].running_mean.double(), net[0].running_var.double(), 1.0 / n_l)
    x1 = x.clone().requires_grad_(True)
    m1, v1 = bns_torch.channel_stats(x1)
    ((mse(m1, net[0].running_mean) + mse(v1, net[0].running_var)) / n_l).backward()
    scale = float(x1.grad.abs().max()) + 1e-30
    assert float((x1.grad.double() - g).abs().max()) <= 2e-5 * scale
