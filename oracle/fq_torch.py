"""Op-by-op restatement of the reference fake-quant arithmetic on torch CPU tensors.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows; paths are relative to ``/root/reference``.  Each
torch call below is one fp32 rounding, exactly as in the reference where every
arithmetic step is its own ATen launch (SURVEY.md section 8(a')).

Nothing here is tuned: this is also the ``kind="port"`` CPU baseline that
``bench.py`` times, so it must cost what the reference costs on host cores
(six full-tensor passes per fake-quant).
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- params
def _levels(k: int) -> float:
    return float(2 ** k - 1)


def _half(k: int) -> float:
    return float(2 ** (k - 1))


def quant_params(k, lo, hi, integral_zero_point=True, signed=True):
    """scale / zero-point from a saturation range.

    quantization_utils/quant_utils.py:117-128 (and the DSG twin :248-259, which
    is the same formula).  ``n / tensor`` in the reference is
    ``Tensor.__rdiv__`` = ``tensor.reciprocal() * n`` -- NOT an IEEE ``n / r``.
    """
    span = torch.clamp(hi - lo, min=1e-8)
    scale = span.reciprocal() * _levels(k)
    zp = scale * lo
    if integral_zero_point:
        zp = zp.round()
    if signed:
        zp = zp + _half(k)
    return scale, zp


def _per_row(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """quant_utils.py:70-76 / :93-99: scale and zero-point index dim 0."""
    if like.dim() == 4:
        return t.reshape(-1, 1, 1, 1)
    if like.dim() == 2:
        return t.reshape(-1, 1)
    return t


def quantize(x, scale, zp):
    """quant_utils.py:81 -- ``round(scale * x - zp)``: mul, sub, round-half-even."""
    s = _per_row(scale, x)
    z = _per_row(zp, x)
    t = s * x
    t = t - z
    return torch.round(t)


def saturate(q, k):
    """quant_utils.py:151-152 -- codes live in [-2^(k-1), 2^(k-1)-1]."""
    h = 2 ** (k - 1)
    return torch.clamp(q, -h, h - 1)


def dequantize(q, scale, zp):
    """quant_utils.py:104 -- ``(q + zp) / scale`` with a true IEEE division."""
    s = _per_row(scale, q)
    z = _per_row(zp, q)
    t = q + z
    return t / s


def codes(x, k, lo, hi):
    """Integer codes (integer-valued fp32) the reference forms at quant_utils.py:148-152."""
    scale, zp = quant_params(k, lo, hi)
    return saturate(quantize(x, scale, zp), k)


def fake_quant(x, k, lo, hi):
    """AsymmetricQuantFunction.forward, quant_utils.py:138-157."""
    scale, zp = quant_params(k, lo, hi)
    q = saturate(quantize(x, scale, zp), k)
    return dequantize(q, scale, zp)


def codes_symmetric(x, k, lo, hi):
    """quant_utils.py:277-282 -- ``round(scale * x)``, zero-point unused."""
    scale, _ = quant_params(k, lo, hi)
    s = _per_row(scale, x)
    return saturate(torch.round(s * x), k)


def fake_quant_symmetric(x, k, lo, hi):
    """SymmetricQuantFunction_DSG.forward, quant_utils.py:268-286."""
    scale, _ = quant_params(k, lo, hi)
    s = _per_row(scale, x)
    q = saturate(torch.round(s * x), k)
    return q / s


class _STE(torch.autograd.Function):
    """Identity straight-through backward, quant_utils.py:159-161 / :288-290."""

    @staticmethod
    def forward(ctx, x, k, lo, hi, symmetric):
        fn = fake_quant_symmetric if symmetric else fake_quant
        return fn(x, k, lo, hi)

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None


def fake_quant_ste(x, k, lo, hi, symmetric=False):
    return _STE.apply(x, k, lo, hi, symmetric)


# --------------------------------------------------------------------------- ranges
def ema_step(state, sample, beta, beta_t_new):
    """One bias-corrected running-range step, quant_modules.py:88-89.

    The corrected value is written BACK into the state (not a textbook EMA).
    ``beta_t_new`` is beta_t after the multiply at :87.
    """
    a = state * beta
    b = sample * (1 - beta)
    c = a + b
    return c / (1 - beta_t_new)


def symmetric_bounds(lo, hi):
    """quant_modules.py:369-374 -- range = +-max(|min|, |max|)."""
    m = torch.maximum(lo.abs(), hi.abs())
    # the reference's first branch is taken only on strict '>', both give +-m
    return -m, m


def row_minmax(w):
    """quant_modules.py:220-222 / :271-273 -- per-output-row min and max."""
    rows = w.detach().contiguous().view(w.shape[0], -1)
    return rows.min(dim=1).values, rows.max(dim=1).values


def row_absmax(w):
    """quant_modules.py:426-427 / :473-474."""
    rows = w.detach().contiguous().view(w.shape[0], -1)
    m = rows.abs().max(dim=1).values
    return -m, m


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    """quant_utils.py:26-33."""
    d = (pred - tgt).abs().pow(p)
    if reduction == "none":
        return d.sum(1).mean()
    return d.mean()


def mse_range_search(x, k, lo, hi, steps=80, p=2.4):
    """Clip-ratio search of QuantAct_MSE, quant_modules.py:160-174.

    Returns the (lo, hi) pair with the smallest L_p score; first strict
    improvement wins, as in the reference loop.
    """
    best = 1e10
    keep = (lo, hi)
    for i in range(steps):
        f = 1.0 - (i * 0.01)
        cand_lo, cand_hi = lo * f, hi * f
        score = lp_loss(x, fake_quant(x, k, cand_lo, cand_hi), p=p, reduction="all")
        if score < best:
            best = score
            keep = (cand_lo, cand_hi)
    return keep


# --------------------------------------------------------------------------- modules
class OracleQuantAct(nn.Module):
    """QuantAct, quant_modules.py:32-96 (``symmetric=True`` gives QuantAct_DSG :315-386)."""

    symmetric = False

    def __init__(self, activation_bit, full_precision_flag=False, running_stat=True, beta=0.9):
        super().__init__()
        self.activation_bit = activation_bit
        self.full_precision_flag = full_precision_flag
        self.running_stat = running_stat
        self.register_buffer("x_min", torch.zeros(1))
        self.register_buffer("x_max", torch.zeros(1))
        self.register_buffer("beta", torch.tensor([beta], dtype=torch.float32))
        self.register_buffer("beta_t", torch.ones(1))

    def fix(self):
        self.running_stat = False

    def unfix(self):
        self.running_stat = True

    def observe(self, x):
        lo = x.detach().min()
        hi = x.detach().max()
        if self.symmetric:
            lo, hi = symmetric_bounds(lo, hi)
        self.beta_t = self.beta_t * self.beta
        self.x_min = ema_step(self.x_min, lo, self.beta, self.beta_t)
        self.x_max = ema_step(self.x_max, hi, self.beta, self.beta_t)

    def forward(self, x):
        if self.running_stat:
            self.observe(x)
        if self.full_precision_flag:
            return x
        return fake_quant_ste(x, self.activation_bit, self.x_min, self.x_max, self.symmetric)


class OracleQuantActSym(OracleQuantAct):
    symmetric = True


class OracleQuantActMSE(OracleQuantAct):
    """QuantAct_MSE, quant_modules.py:98-186: searched range, EMA without bias correction."""

    def observe(self, x):
        xd = x.detach().clone()
        lo, hi = mse_range_search(xd, self.activation_bit, xd.min(), xd.max())
        self.beta_t = self.beta_t * self.beta
        self.x_min = self.x_min * self.beta + lo * (1 - self.beta)
        self.x_max = self.x_max * self.beta + hi * (1 - self.beta)


class _OracleQuantWeight(nn.Module):
    symmetric = False

    def __init__(self, weight_bit, full_precision_flag=False):
        super().__init__()
        self.weight_bit = weight_bit
        self.full_precision_flag = full_precision_flag

    def _adopt(self, src):
        self.weight = nn.Parameter(src.weight.data.clone())
        self.bias = None if src.bias is None else nn.Parameter(src.bias.data.clone())

    def quantized_weight(self):
        if self.full_precision_flag:
            return self.weight
        lo, hi = (row_absmax if self.symmetric else row_minmax)(self.weight)
        return fake_quant_ste(self.weight, self.weight_bit, lo, hi, self.symmetric)


class OracleQuantConv2d(_OracleQuantWeight):
    """Quant_Conv2d, quant_modules.py:235-281 (symmetric: QuantConv2d_DSG :436-481)."""

    def set_param(self, conv):
        for a in ("in_channels", "out_channels", "kernel_size", "stride", "padding", "dilation", "groups"):
            setattr(self, a, getattr(conv, a))
        self._adopt(conv)

    def forward(self, x):
        return F.conv2d(x, self.quantized_weight(), self.bias, self.stride, self.padding,
                        self.dilation, self.groups)


class OracleQuantLinear(_OracleQuantWeight):
    """Quant_Linear, quant_modules.py:188-232 (symmetric: QuantLinear_DSG :389-433)."""

    def set_param(self, linear):
        self.in_features = linear.in_features
        self.out_features = linear.out_features
        self._adopt(linear)

    def forward(self, x):
        return F.linear(x, self.quantized_weight(), self.bias)


class OracleQuantConv2dSym(OracleQuantConv2d):
    symmetric = True


class OracleQuantLinearSym(OracleQuantLinear):
    symmetric = True
