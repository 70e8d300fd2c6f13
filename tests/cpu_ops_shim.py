"""TEST-ONLY stand-in for ``ood_dfq_b200.ops`` on CPU tensors, backed by the oracle.

The mirror modules (ood_dfq_b200/quantization_utils) have no CPU path: every tensor operation is a kernel launch.
Their HOST logic -- which autograd Function runs when, how the range buffers are updated in place, when the
WeightBank re-quantises, what ``fix`` / ``full_precision_flag`` / deepcopy do -- is plain Python and can be exercised
without a GPU by swapping the handful of launch wrappers the modules call for oracle arithmetic with the same
contracts (in-place state updates included).  ``installed()`` is a context manager used by
tests/test_mirror_host_logic.py; nothing in the package can reach this file.
"""
import contextlib

import torch

from oracle import fq_torch


def quant_params(k, lo, hi):
    return fq_torch.quant_params(k, lo, hi)


def fake_quant(x, k, lo, hi, symmetric=False, codes=False, relu_first=False):
    x = torch.relu(x) if relu_first else x
    fn = fq_torch.fake_quant_symmetric if symmetric else fq_torch.fake_quant
    y = fn(x.detach(), k, lo, hi)
    if codes:
        cfn = fq_torch.codes_symmetric if symmetric else fq_torch.codes
        return y, cfn(x.detach(), k, lo, hi).to(torch.int8)
    return y


def elementwise(x, p0, p1, k, mode, symmetric=False, params_given=False, out=None, codes=False, relu_first=False):
    from ood_dfq_b200 import _native as N
    assert params_given and not codes and not relu_first, "the shim covers the helper calls of quant_utils only"
    if mode == N.MODE_QUANTIZE:
        y = torch.round(fq_torch._per_row(p0, x) * x) if symmetric else fq_torch.quantize(x, p0, p1)
    elif mode == N.MODE_DEQUANTIZE:
        y = x / fq_torch._per_row(p0, x) if symmetric else fq_torch.dequantize(x, p0, p1)
    else:
        raise AssertionError(mode)
    if out is not None:
        out.copy_(y)
        return out
    return y


def minmax(x):
    return torch.stack([x.detach().min(), x.detach().max()])


def act_calib_forward(x, k, x_min, x_max, beta, beta_t, symmetric=False, quantize=True, codes=False, onchip=True):
    """In place on the buffers, like the kernel's last CTA (quant_modules.py:87-89 arithmetic)."""
    lo, hi = x.detach().min(), x.detach().max()
    if symmetric:
        lo, hi = fq_torch.symmetric_bounds(lo, hi)
    bt = beta_t * beta
    new_min = fq_torch.ema_step(x_min, lo, beta, bt)
    new_max = fq_torch.ema_step(x_max, hi, beta, bt)
    x_min.copy_(new_min)
    x_max.copy_(new_max)
    beta_t.copy_(bt)
    if not quantize:
        return None
    return fake_quant(x, k, x_min, x_max, symmetric=symmetric, codes=codes)


def act_calib_stats_forward(x, k, x_min, x_max, beta, beta_t, sums=None, onchip=True):
    """Range update + fake-quant + raw per-channel fp64 sums of the same input (csrc/fq_calib.cu contract)."""
    y = act_calib_forward(x, k, x_min, x_max, beta, beta_t)
    return y, bn_stats_forward(x, None, sums=sums)


def weight_fq_multi(weights, ks, symmetric, outs=None, want_range=False, want_codes=False):
    res = []
    for i, w in enumerate(weights):
        wd = w.detach()
        rows = wd.contiguous().view(wd.shape[0], -1)
        if symmetric[i]:
            hi = rows.abs().max(dim=1).values
            lo = -hi
            wq = fq_torch.fake_quant_symmetric(wd, ks[i], lo, hi)
        else:
            lo, hi = rows.min(dim=1).values, rows.max(dim=1).values
            wq = fq_torch.fake_quant(wd, ks[i], lo, hi)
        if outs is not None:
            outs[i].copy_(wq)
            wq = outs[i]
        res.append({"wq": wq, "lo": lo if want_range else None, "hi": hi if want_range else None, "codes": None})
    return res


def act_mse_search(x, k, x_min, x_max, beta, beta_t, cur_min=None, cur_max=None, steps=80, step=0.01, p=2.4,
                   debug=False):
    xd = x.detach().clone()
    lo, hi = xd.min(), xd.max()
    if cur_min is not None:
        cur_min.copy_(lo.reshape(1))
        cur_max.copy_(hi.reshape(1))
    keep_lo, keep_hi = fq_torch.mse_range_search(xd, k, lo, hi, steps=steps, p=p)
    beta_t.copy_(beta_t * beta)
    x_min.copy_(x_min * beta + keep_lo * (1 - beta))
    x_max.copy_(x_max * beta + keep_hi * (1 - beta))


# ---- BN-statistics loss (csrc/bn_stats.cu contracts: shifted fp64 sums, packed layer blocks [S1 | S2]) ----------------
def bn_stats_forward(x, shift=None, sums=None, fq=None):
    c = x.shape[1]
    sh = torch.zeros(c, dtype=torch.float64) if shift is None else shift.detach().double()
    d = x.detach().double() - sh.view(1, c, 1, 1)
    block = torch.cat([d.sum([0, 2, 3]), (d * d).sum([0, 2, 3])])
    if sums is None:
        sums = block
    else:
        sums.copy_(block)
    return (sums, fake_quant(x, fq[0], fq[1], fq[2])) if fq is not None else sums


def bn_stats_finalize(sums, shift, count):
    c = sums.numel() // 2
    m1 = sums[:c] / count
    mean = m1 if shift is None else shift.double() + m1
    return mean.float(), (sums[c:] / count - m1 * m1).float()


def bns_loss(sums, shift, run_mean, run_var, ch_off, counts):
    n_layers, ctot = len(counts), ch_off[-1]
    mean, var = torch.empty(ctot, dtype=torch.float64), torch.empty(ctot, dtype=torch.float64)
    gmean, gvar = torch.empty(ctot, dtype=torch.float64), torch.empty(ctot, dtype=torch.float64)
    t_mean = t_var = 0.0
    for l in range(n_layers):
        a, b = ch_off[l], ch_off[l + 1]
        c = b - a
        block = sums[2 * a: 2 * b]
        m1 = block[:c] / counts[l]
        mean[a:b] = (shift[a:b].double() if shift is not None else 0.0) + m1
        var[a:b] = block[c:] / counts[l] - m1 * m1
        dm, dv = mean[a:b] - run_mean[a:b].double(), var[a:b] - run_var[a:b].double()
        t_mean = t_mean + (dm * dm).mean()
        t_var = t_var + (dv * dv).mean()
        gmean[a:b] = 2.0 * dm / (c * n_layers)
        gvar[a:b] = 2.0 * dv / (c * n_layers)
    loss3 = torch.stack([(t_mean + t_var) / n_layers, t_mean / n_layers, t_var / n_layers]).float()
    return loss3, mean.float(), var.float(), gmean.float(), gvar.float()


def bn_stats_backward(x, grad_in, mean, gmean, gvar, count, gscale=None, out=None):
    c = x.shape[1]
    g = 1.0 if gscale is None else gscale.double().reshape(())
    v = lambda t: t.double().view(1, c, 1, 1)
    r = g * (v(gmean) / count + v(gvar) * 2.0 * (x.detach().double() - v(mean)) / count)
    if grad_in is not None:
        r = r + grad_in.double()
    r = r.float().contiguous(memory_format=torch.channels_last if x.is_contiguous(memory_format=torch.channels_last)
                             and not x.is_contiguous() else torch.contiguous_format)
    if out is not None:
        out.copy_(r)
        return out
    return r


PATCHED = ("quant_params", "fake_quant", "elementwise", "minmax", "act_calib_forward", "act_calib_stats_forward", "weight_fq_multi", "act_mse_search",
           "bn_stats_forward", "bn_stats_finalize", "bns_loss", "bn_stats_backward")


@contextlib.contextmanager
def installed():
    from ood_dfq_b200 import ops
    from ood_dfq_b200.quantization_utils.quant_modules import WeightBank
    saved = {name: getattr(ops, name) for name in PATCHED}
    try:
        for name in PATCHED:
            setattr(ops, name, globals()[name])
        WeightBank.invalidate()
        yield
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
        WeightBank.invalidate()
