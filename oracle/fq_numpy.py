"""numpy-float32 restatement of the fake-quant arithmetic, one rounding per line.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  This is the listing of
SURVEY.md section 8(a') turned into code, independent of torch, so the golden vectors
are checked by two different restatements.  All values are np.float32; every
binary op on float32 arrays is a single correctly-rounded IEEE operation and
``np.rint`` rounds half to even.  Paths relative to ``/root/reference``.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def quant_params(k, lo, hi):
    """quantization_utils/quant_utils.py:117-128."""
    lo = np.asarray(lo, dtype=f32)
    hi = np.asarray(hi, dtype=f32)
    r = hi - lo
    r = np.maximum(r, f32(1e-8))
    # NaN in, NaN out (torch.clamp keeps NaN; np.maximum propagates it too)
    inv = f32(1.0) / r
    scale = inv * f32(2 ** k - 1)
    z = scale * lo
    z = np.rint(z)
    zp = z + f32(2 ** (k - 1))
    return scale.astype(f32), zp.astype(f32)


def _rows(p, x):
    p = np.asarray(p, dtype=f32).reshape(-1)
    if x.ndim in (2, 4) :
        return p.reshape((-1,) + (1,) * (x.ndim - 1))
    return p


def codes(x, k, lo, hi, symmetric=False):
    """quant_utils.py:81 + :151-152 (symmetric: :212, :281-282)."""
    x = np.asarray(x, dtype=f32)
    scale, zp = quant_params(k, lo, hi)
    a = _rows(scale, x) * x
    b = a if symmetric else a - _rows(zp, x)
    q = np.rint(b)
    h = f32(2 ** (k - 1))
    with np.errstate(invalid="ignore"):
        q = np.minimum(np.maximum(q, -h), h - f32(1))
    return q.astype(f32)


def fake_quant(x, k, lo, hi, symmetric=False):
    """quant_utils.py:148-157 (symmetric: :277-286)."""
    x = np.asarray(x, dtype=f32)
    scale, zp = quant_params(k, lo, hi)
    q = codes(x, k, lo, hi, symmetric)
    c = q if symmetric else q + _rows(zp, x)
    return (c / _rows(scale, x)).astype(f32)


def ema_step(state, sample, beta, beta_t_new):
    """quant_modules.py:88-89 with beta_t already multiplied (:87)."""
    state, sample, beta, beta_t_new = (f32(v) for v in (state, sample, beta, beta_t_new))
    omb = f32(1.0) - beta
    t1 = state * beta
    t2 = sample * omb
    t3 = t1 + t2
    d = f32(1.0) - beta_t_new
    return f32(t3 / d)


def range_update(x_min, x_max, beta, beta_t, data_min, data_max, symmetric=False):
    """Full calibrating step of QuantAct (quant_modules.py:80-89; DSG :365-380)."""
    data_min, data_max = f32(data_min), f32(data_max)
    if symmetric:
        m = max(abs(data_min), abs(data_max))
        data_min, data_max = f32(-m), f32(m)
    beta_t = f32(f32(beta_t) * f32(beta))
    return (ema_step(x_min, data_min, beta, beta_t),
            ema_step(x_max, data_max, beta, beta_t),
            beta_t)
