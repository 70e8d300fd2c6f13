"""Pin the oracle: replay the reference-generated golden vectors through both restatements.

CPU only.  Bit-exact (int32 views) for codes, dequantised values and range state.
"""
import numpy as np
import pytest
import torch

from conftest import bits
from oracle import bns_torch, fq_numpy, fq_torch

T = torch.from_numpy
FROZEN_CASES = [f"{t}_k{k}" for t in ("hw49", "hw16", "flat") for k in (2, 3, 4, 8)] + \
               [f"signed_k{k}" for k in (2, 4, 8)] + ["ties_k4", "degen_k4"]


@pytest.mark.parametrize("case", FROZEN_CASES)
def test_frozen_activation(golden, case):
    g = golden("act_frozen")
    k = int(case.rsplit("_k", 1)[1])
    x, lo, hi = (g[f"{case}_{n}"] for n in ("x", "lo", "hi"))
    s, z = fq_torch.quant_params(k, T(lo), T(hi))
    assert np.array_equal(bits(s.numpy()), bits(g[f"{case}_scale"]))
    assert np.array_equal(bits(z.numpy()), bits(g[f"{case}_zp"]))
    assert np.array_equal(bits(fq_torch.codes(T(x), k, T(lo), T(hi)).numpy()), bits(g[f"{case}_codes"]))
    assert np.array_equal(bits(fq_torch.fake_quant(T(x), k, T(lo), T(hi)).numpy()), bits(g[f"{case}_y"]))
    # the independent numpy restatement
    s2, z2 = fq_numpy.quant_params(k, lo, hi)
    assert np.array_equal(bits(s2), bits(g[f"{case}_scale"]))
    assert np.array_equal(bits(z2), bits(g[f"{case}_zp"]))
    assert np.array_equal(bits(fq_numpy.codes(x, k, lo, hi)), bits(g[f"{case}_codes"]))
    assert np.array_equal(bits(fq_numpy.fake_quant(x, k, lo, hi)), bits(g[f"{case}_y"]))
    h = 2 ** (k - 1)
    c = g[f"{case}_codes"]
    assert c.min() >= -h and c.max() <= h - 1 and np.array_equal(c, np.rint(c))


@pytest.mark.parametrize("tag,cls", [("asym", fq_torch.OracleQuantAct), ("sym", fq_torch.OracleQuantActSym)])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_calibrating_sequence(golden, tag, cls, k):
    g = golden("act_calib")
    p = f"{tag}_k{k}_"
    m = cls(k)
    st = (np.float32(0), np.float32(0), np.float32(1))
    for step in range(6):
        x = g[p + f"x{step}"]
        if step == 4:
            m.fix()
        if step == 5:
            m.unfix()
        y = m(T(x))
        assert np.array_equal(bits(y.numpy()), bits(g[p + f"y{step}"])), step
        state = np.concatenate([m.x_min.numpy(), m.x_max.numpy(), m.beta_t.numpy()])
        assert np.array_equal(bits(state), g[p + f"state_bits{step}"]), step
        if step != 4:   # numpy restatement of the recurrence
            st = fq_numpy.range_update(st[0], st[1], g[p + "beta"][0], st[2], x.min(), x.max(), symmetric=(tag == "sym"))
        assert np.array_equal(bits(np.array(st, dtype=np.float32)), g[p + f"state_bits{step}"]), step
        y2 = fq_numpy.fake_quant(x, k, st[0], st[1], symmetric=(tag == "sym"))
        assert np.array_equal(bits(y2), bits(g[p + f"y{step}"])), step


def test_full_precision_passthrough(golden):
    g = golden("act_calib")
    m = fq_torch.OracleQuantAct(4, full_precision_flag=True)
    x = T(g["fp_x"])
    assert m(x) is x
    state = np.array([m.x_min.item(), m.x_max.item(), m.beta_t.item()], dtype=np.float32)
    assert np.array_equal(bits(state), bits(g["fp_state"]))


@pytest.mark.parametrize("tag", ["c3x3", "c1x1", "c7x7", "wide"])
@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("sym", [False, True])
def test_conv_weights(golden, tag, k, sym):
    g = golden("weights")
    p = f"{tag}_k{k}_{'sym' if sym else 'asym'}_"
    w = T(g[p + "w"])
    lo, hi = (fq_torch.row_absmax if sym else fq_torch.row_minmax)(w)
    assert np.array_equal(bits(lo.numpy()), bits(g[p + "lo"]))
    assert np.array_equal(bits(hi.numpy()), bits(g[p + "hi"]))
    code_fn = fq_torch.codes_symmetric if sym else fq_torch.codes
    fq_fn = fq_torch.fake_quant_symmetric if sym else fq_torch.fake_quant
    assert np.array_equal(bits(code_fn(w, k, lo, hi).numpy()), bits(g[p + "codes"]))
    assert np.array_equal(bits(fq_fn(w, k, lo, hi).numpy()), bits(g[p + "wq"]))
    assert np.array_equal(bits(fq_numpy.fake_quant(g[p + "w"], k, g[p + "lo"], g[p + "hi"], symmetric=sym)),
                          bits(g[p + "wq"]))
    # module-level: forward output and STE weight gradient
    cls = fq_torch.OracleQuantConv2dSym if sym else fq_torch.OracleQuantConv2d
    ksz = w.shape[2]
    conv = torch.nn.Conv2d(w.shape[1], w.shape[0], ksz, padding=ksz // 2, bias=(p + "bias") in g)
    with torch.no_grad():
        conv.weight.copy_(w)
        if conv.bias is not None:
            conv.bias.copy_(T(g[p + "bias"]))
    qm = cls(k)
    qm.set_param(conv)
    out = qm(T(g[p + "x"]))
    out.square().sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), g[p + "out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(qm.weight.grad.numpy(), g[p + "wgrad"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("k", [2, 4, 8])
@pytest.mark.parametrize("sym", [False, True])
def test_linear_weights(golden, k, sym):
    g = golden("weights")
    p = f"lin_k{k}_{'sym' if sym else 'asym'}_"
    lin = torch.nn.Linear(64, 10)
    with torch.no_grad():
        lin.weight.copy_(T(g[p + "w"]))
        lin.bias.copy_(T(g[p + "bias"]))
    qm = (fq_torch.OracleQuantLinearSym if sym else fq_torch.OracleQuantLinear)(k)
    qm.set_param(lin)
    out = qm(T(g[p + "x"]))
    out.square().sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), g[p + "out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(qm.weight.grad.numpy(), g[p + "wgrad"], rtol=1e-5, atol=1e-5)


def test_constant_rows(golden):
    g = golden("weights")
    w = T(g["constrow_w"])
    lo, hi = fq_torch.row_minmax(w)
    assert np.array_equal(bits(fq_torch.fake_quant(w, 4, lo, hi).numpy()), bits(g["constrow_wq"]))


def test_mse_searched_range(golden):
    g = golden("act_mse")
    m = fq_torch.OracleQuantActMSE(4)
    for step in range(2):
        y = m(T(g[f"x{step}"]))
        state = np.array([m.x_min.item(), m.x_max.item(), m.beta_t.item()], dtype=np.float32)
        assert np.array_equal(bits(state), bits(g[f"state{step}"]))
        assert np.array_equal(bits(y.numpy()), bits(g[f"y{step}"]))


class _TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.c1, self.b1 = nn.Conv2d(3, 6, 3, padding=1, bias=False), nn.BatchNorm2d(6)
        self.c2, self.b2 = nn.Conv2d(6, 10, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(10)
        self.c3, self.b3 = nn.Conv2d(10, 4, 1, bias=False), nn.BatchNorm2d(4)

    def forward(self, x):
        x = torch.relu(self.b1(self.c1(x)))
        x = torch.relu(self.b2(self.c2(x)))
        return self.b3(self.c3(x))


def tiny_net(g):
    net = _TinyNet().eval()
    net.load_state_dict({k[len("param_"):]: T(v) for k, v in g.items() if k.startswith("param_")})
    return net


@pytest.mark.parametrize("flavour", ["trainer", "distill"])
def test_bns_loss_and_grad(golden, flavour):
    g = golden("bns")
    net = tiny_net(g)
    tap = bns_torch.StatTap(net)
    x = T(g["x"]).clone().requires_grad_(True)
    net(x)
    for i in range(3):
        np.testing.assert_allclose(tap.means[i].detach().numpy(), g[f"mean{i}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(tap.vars[i].detach().numpy(), g[f"var{i}"], rtol=1e-6, atol=1e-7)
    loss = tap.loss(flavour)
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g[f"{flavour}_loss"], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g[f"{flavour}_xgrad"], rtol=1e-5, atol=1e-9)


def test_bns_closed_form_gradient(golden):
    g = golden("bns")
    x = T(g["stat_off0_x"]).clone().requires_grad_(True)
    rm = torch.linspace(-0.2, 0.3, x.shape[1])
    rv = torch.linspace(0.6, 1.4, x.shape[1])
    mean, var = bns_torch.channel_stats(x)
    loss = torch.nn.functional.mse_loss(mean, rm) + torch.nn.functional.mse_loss(var, rv)
    loss.backward()
    np.testing.assert_allclose(bns_torch.bns_input_grad(x.detach(), rm, rv).numpy(), x.grad.numpy(), rtol=2e-5, atol=1e-9)


# ------------------------------------------------------------------ the plain-C restatement (gcc)
from oracle import fq_c  # noqa: E402


@pytest.mark.parametrize("case", FROZEN_CASES)
def test_c_oracle_frozen(golden, case):
    g = golden("act_frozen")
    k = int(case.rsplit("_k", 1)[1])
    x, lo, hi = (g[f"{case}_{n}"] for n in ("x", "lo", "hi"))
    s, z = fq_c.params(k, lo[0], hi[0])
    assert bits(np.array([s, z])).tolist() == bits(np.array([g[f"{case}_scale"][0], g[f"{case}_zp"][0]])).tolist()
    y, codes = fq_c.fake_quant(x, k, lo, hi)
    assert np.array_equal(bits(y), bits(g[f"{case}_y"]))
    assert np.array_equal(bits(codes), bits(g[f"{case}_codes"]))


@pytest.mark.parametrize("tag", ["asym", "sym"])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_c_oracle_calibrating_sequence(golden, tag, k):
    g = golden("act_calib")
    p = f"{tag}_k{k}_"
    st = np.array([0, 0, 1], dtype=np.float32)
    for step in range(6):
        x = g[p + f"x{step}"]
        if step != 4:
            mn, mx = fq_c.minmax(x)
            st = fq_c.range_update(st, g[p + "beta"][0], mn, mx, symmetric=(tag == "sym"))
        assert np.array_equal(bits(st), g[p + f"state_bits{step}"]), step
        y, _ = fq_c.fake_quant(x, k, st[0], st[1], symmetric=(tag == "sym"))
        assert np.array_equal(bits(y), bits(g[p + f"y{step}"])), step


@pytest.mark.parametrize("tag", ["c3x3", "c1x1", "c7x7", "wide"])
@pytest.mark.parametrize("sym", [False, True])
def test_c_oracle_weights(golden, tag, sym):
    g = golden("weights")
    for k in (2, 4, 8):
        p = f"{tag}_k{k}_{'sym' if sym else 'asym'}_"
        lo, hi = fq_c.row_ranges(g[p + "w"], symmetric=sym)
        assert np.array_equal(bits(lo), bits(g[p + "lo"])) and np.array_equal(bits(hi), bits(g[p + "hi"]))
        y, codes = fq_c.fake_quant(g[p + "w"], k, lo, hi, symmetric=sym)
        assert np.array_equal(bits(y), bits(g[p + "wq"]))
        assert np.array_equal(bits(codes), bits(g[p + "codes"]))


def test_c_oracle_channel_stats(golden):
    g = golden("bns")
    for tag in ("off0", "off10", "off100"):
        mean, var = fq_c.channel_stats(g[f"stat_{tag}_x"])
        np.testing.assert_allclose(mean, g[f"stat_{tag}_mean"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(var, g[f"stat_{tag}_var64"], rtol=1e-9)


def test_c_mse_search_matches_the_reference(golden):
    """The plain-C restatement of QuantAct_MSE's search reproduces the reference-generated state and output."""
    g = golden("act_mse")
    st = np.array([0.0, 0.0, 1.0], np.float32)
    for step in range(2):
        st, scores, keep = fq_c.mse_search(g[f"x{step}"], 4, st)
        assert 0 <= keep < 80 and np.all(np.isfinite(scores))
        np.testing.assert_allclose(st, g[f"state{step}"], rtol=1e-6)
        y, _ = fq_c.fake_quant(g[f"x{step}"], 4, st[0], st[1])
        np.testing.assert_allclose(y, g[f"y{step}"], rtol=1e-5, atol=1e-6)
