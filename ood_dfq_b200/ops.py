"""Tensor-level entry points: argument checks, output allocation, one C-ABI call each.

Every function takes CUDA fp32 tensors and launches hand-written sm_100a kernels on
the current stream through ``include/oodfq_b200.h``.  Anything else (CPU tensors,
other dtypes) raises -- there is deliberately no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _native as N

_workspaces: dict = {}

# bench.py sets this to a list to time the streaming launches with CUDA events on the launching
# stream: entries are (kernel family, start_event, end_event, algorithmic_bytes).
PROFILE = None


class _Timed:
    """``with _Timed(name, nbytes):`` brackets one launch with events when PROFILE is armed."""

    __slots__ = ("name", "nbytes", "ev0")

    def __init__(self, name, nbytes):
        self.name, self.nbytes, self.ev0 = name, nbytes, None

    def __enter__(self):
        if PROFILE is not None:
            self.ev0 = torch.cuda.Event(enable_timing=True)
            self.ev0.record()

    def __exit__(self, *exc):
        if self.ev0 is not None and exc[0] is None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            PROFILE.append((self.name, self.ev0, ev1, self.nbytes))
        return False


def _need(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"ood_dfq_b200: {name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"ood_dfq_b200: {name} is on {t.device}; this path only runs on CUDA "
                           "(sm_100a kernels, no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"ood_dfq_b200: {name} must be {dtype}, got {t.dtype}")


def _dense(t: torch.Tensor) -> torch.Tensor:
    """A tensor whose storage can be walked flat (NCHW- or NHWC-contiguous); else a contiguous copy."""
    if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)):
        return t
    return t.contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _written(t: torch.Tensor) -> torch.Tensor:
    """Tell torch that a kernel wrote into a caller-provided tensor through its raw pointer: an in-place op on an
    empty slice bumps the (shared) version counter without launching anything.  Autograd's saved-tensor checks and the
    caches keyed on ``_version`` (the space-to-depth stem input, fusion.py) then see the write."""
    if t.dim() > 0:
        with torch.no_grad():                 # only the counter matters: no autograd node for this no-op
            t[:0].zero_()
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------- deferred parameter-gradient folds
_DEFER_KEEP = None          # while folds are deferred: the dwdb buffers handed out, kept alive until the flush
_FOLD_ARENAS = {}


class deferred_folds:
    """``with ops.deferred_folds(device):`` -- the BatchNorm parameter-gradient reductions launched inside (fused BN,
    residual-tail and stem backward with parameter gradients) leave their per-CTA partials in an arena and are folded
    by ONE launch when the block ends, instead of one small launch behind each of them (34-38 per QAT iteration).

    Until the block ends the dW / dB tensors those backwards returned hold garbage: use it only around a sweep whose
    results nothing reads before the block is left (``step.QATStep`` wraps each ``autograd.grad`` over the parameters).
    Not re-entrant; bit-identical results."""

    def __init__(self, device, arena_mb=96):
        self.device, self.bytes = torch.device(device), int(arena_mb) << 20

    def __enter__(self):
        global _DEFER_KEEP
        if self.device.type != "cuda":
            raise RuntimeError("ood_dfq_b200: deferred_folds only exists on CUDA devices")
        key = (self.device.index if self.device.index is not None else torch.cuda.current_device(), self.bytes)
        arena = _FOLD_ARENAS.get(key)
        if arena is None:
            arena = _FOLD_ARENAS[key] = torch.empty(self.bytes, dtype=torch.uint8, device=self.device)
        N.check(N.load().oodfq_defer_folds_begin(arena.data_ptr(), self.bytes), "defer_folds_begin")
        _DEFER_KEEP = []
        return self

    def __exit__(self, *exc):
        global _DEFER_KEEP
        rc = N.load().oodfq_defer_folds_end(_stream(self.device))
        _DEFER_KEEP = None
        if exc[0] is None:
            N.check(rc, "defer_folds_end")
        return False


def _keep_until_flush(t):
    if _DEFER_KEEP is not None and t is not None:
        _DEFER_KEEP.append(t)
    return t


def workspace(device) -> torch.Tensor:
    """Zero-initialised scratch for the reducing kernels, one per (device, stream)."""
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           _stream(device))
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(N.load().oodfq_workspace_bytes()), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ----------------------------------------------------------------------------- a1
def quant_params(k: int, lo: torch.Tensor, hi: torch.Tensor):
    """(scale, zero_point) of a saturation range; reference quant_utils.py:107-128 / :238-259."""
    _need(lo, "saturation_min")
    _need(hi, "saturation_max")
    if lo.shape != hi.shape:
        lo, hi = torch.broadcast_tensors(lo, hi)
    lo_c, hi_c = lo.contiguous(), hi.contiguous()
    scale = torch.empty_like(lo_c)
    zp = torch.empty_like(lo_c)
    rc = N.load().oodfq_quant_params(lo_c.data_ptr(), hi_c.data_ptr(), scale.data_ptr(), zp.data_ptr(),
                                     lo_c.numel(), int(k), _stream(lo.device))
    N.check(rc, "quant_params")
    return scale, zp


# ----------------------------------------------------------------------------- a2-a5
def _rows_for(x: torch.Tensor, p: torch.Tensor, what: str) -> int:
    n = p.numel()
    if n == 1:
        return 1
    # the reference reshapes scale/zero-point to (-1,1,1,1) or (-1,1): one value per dim-0 row
    if x.dim() not in (2, 4) or n != x.shape[0]:
        raise RuntimeError(f"ood_dfq_b200: {what} has {n} entries; expected 1 or x.shape[0]={x.shape[0]} "
                           f"for a {x.dim()}-D input")
    return n


def elementwise(x, p0, p1, k, mode, symmetric=False, params_given=False, out=None, codes=False, relu_first=False):
    """One launch of oodfq_fq_forward.  p0/p1 = (min, max) or, with params_given, (scale, zero_point)."""
    _need(x, "input")
    _need(p0, "range/scale")
    _need(p1, "range/zero_point")
    rows = _rows_for(x, p0, "range/scale")
    if p1.numel() != p0.numel():
        raise RuntimeError("ood_dfq_b200: mismatched parameter sizes")
    xd = _dense(x) if rows == 1 else x.contiguous()
    if out is None:
        y = torch.empty_like(xd)
    else:
        y = out
        if y.data_ptr() != xd.data_ptr() or y.shape != xd.shape:
            raise RuntimeError("ood_dfq_b200: `out` must be the (dense) input itself for in-place use")
    cd = torch.empty(xd.shape, dtype=torch.int8, device=x.device) if codes else None
    if cd is not None and xd.stride() != cd.stride():
        cd = torch.empty_like(xd, dtype=torch.int8)
    flags = (N.SYMMETRIC if symmetric else 0) | (N.PARAMS_GIVEN if params_given else 0) | \
        (N.RELU_FIRST if relu_first else 0)
    family = "fq_flat_kernel (QuantAct forward, 8 B/elem)" if (mode == N.MODE_FAKEQUANT and rows == 1) else \
        "fq_* helper kernels (linear_quantize / dequantize / per-row ranges, 8 B/elem)"
    with _Timed(family, 8 * xd.numel()):
        rc = N.load().oodfq_fq_forward(xd.data_ptr(), y.data_ptr(), _ptr(cd), xd.numel(),
                                       p0.contiguous().data_ptr(), p1.contiguous().data_ptr(), rows,
                                       int(k), mode, flags, _stream(x.device))
        N.check(rc, "fq_forward")
    if out is not None:
        _written(y)
    return (y, cd) if codes else y


def fake_quant(x, k, lo, hi, symmetric=False, codes=False, relu_first=False):
    """quantise -> clamp -> dequantise with the reference's rounding sequence (quant_utils.py:138-157).

    ``relu_first``: apply ``max(x, 0)`` in the same pass (the ``Sequential(ReLU, QuantAct)`` of main_direct.py:464-465).
    """
    return elementwise(x, lo, hi, k, N.MODE_FAKEQUANT, symmetric=symmetric, codes=codes, relu_first=relu_first)


# ----------------------------------------------------------------------------- a6
def minmax(x: torch.Tensor) -> torch.Tensor:
    """[min, max] of x, NaN-propagating like torch.min / torch.max (quant_modules.py:81-82)."""
    _need(x, "input")
    xd = _dense(x)
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    rc = N.load().oodfq_minmax(xd.data_ptr(), xd.numel(), out.data_ptr(), workspace(x.device).data_ptr(),
                               _stream(x.device))
    N.check(rc, "minmax")
    return out


def act_calib_forward(x, k, x_min, x_max, beta, beta_t, symmetric=False, quantize=True, codes=False, onchip=True):
    """Calibrating QuantAct forward: updates (x_min, x_max, beta_t) in place, returns fake-quantised x.

    quant_modules.py:80-94 (DSG :365-386).  ``quantize=False`` only tracks the range
    (full_precision_flag) and returns None.  ``onchip=False`` forces the two-kernel path (tensors up to 96 MB
    otherwise run as one cooperative kernel that stages x into shared memory with TMA bulk copies and keeps it there;
    same bits either way).  ``onchip="tma"`` is accepted for round-1 callers and means ``True``.
    """
    _need(x, "input")
    for t, nme in ((x_min, "x_min"), (x_max, "x_max"), (beta, "beta"), (beta_t, "beta_t")):
        _need(t, nme)
        if t.numel() != 1 or not t.is_contiguous():
            raise RuntimeError(f"ood_dfq_b200: {nme} must be a contiguous 1-element buffer")
    xd = _dense(x)
    y = torch.empty_like(xd) if quantize else None
    cd = torch.empty_like(xd, dtype=torch.int8) if (codes and quantize) else None
    rc = N.load().oodfq_act_calib_forward(xd.data_ptr(), _ptr(y), _ptr(cd), xd.numel(), x_min.data_ptr(),
                                          x_max.data_ptr(), beta.data_ptr(), beta_t.data_ptr(), int(k),
                                          (N.SYMMETRIC if symmetric else 0) | (0 if onchip else N.NO_ONCHIP),
                                          workspace(x.device).data_ptr(), _stream(x.device))
    N.check(rc, "act_calib_forward")
    return (y, cd) if codes else y


def act_calib_stats_forward(x, k, x_min, x_max, beta, beta_t, sums=None, onchip=True):
    """Calibrating QuantAct forward that ALSO leaves the per-channel sums of its input (north_star (b)): updates
    (x_min, x_max, beta_t) in place, returns ``(y, sums)`` with ``y`` bit-identical to ``act_calib_forward`` and
    ``sums`` fp64 ``[2*C]`` = (sum_c x, sum_c x^2) -- what ``x.mean([0,2,3])`` / ``x.var([0,2,3], unbiased=False)``
    of the BN-statistics hook (trainer_direct.py:388-393) need, from the same read.

    4-D NCHW-contiguous or channels_last input, k <= 8.  Up to 96 MB one cooperative kernel (x crosses HBM once);
    larger tensors: range reduction + one quantising / accumulating pass.  ``onchip=False`` forces the latter.
    """
    _need(x, "input")
    for t, nme in ((x_min, "x_min"), (x_max, "x_max"), (beta, "beta"), (beta_t, "beta_t")):
        _need(t, nme)
        if t.numel() != 1 or not t.is_contiguous():
            raise RuntimeError(f"ood_dfq_b200: {nme} must be a contiguous 1-element buffer")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    if sums is None:
        sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
    elif sums.dtype != torch.float64 or sums.numel() != 2 * c or not sums.is_contiguous():
        raise RuntimeError("ood_dfq_b200: sums must be a contiguous float64 [2*C] tensor")
    y = torch.empty_like(xc)
    with _Timed("act_calib_stats (single-pass calibrating QuantAct + channel sums, 8 B/elem on chip)", 8 * xc.numel()):
        rc = N.load().oodfq_act_calib_stats_forward(xc.data_ptr(), y.data_ptr(), n, c, hw, x_min.data_ptr(),
                                                    x_max.data_ptr(), beta.data_ptr(), beta_t.data_ptr(), int(k),
                                                    (N.BN_NHWC if nhwc else 0) | (0 if onchip else N.NO_ONCHIP),
                                                    sums.data_ptr(), workspace(x.device).data_ptr(), _stream(x.device))
        N.check(rc, "act_calib_stats_forward")
    return y, sums


# ----------------------------------------------------------------------------- a7 / a8
def weight_fq_multi(weights: Sequence[torch.Tensor], ks: Sequence[int], symmetric: Sequence[bool],
                    outs: Optional[Sequence[torch.Tensor]] = None, want_range=False, want_codes=False):
    """Per-output-row min/max + fake-quant of many weight tensors in ONE launch.

    quant_modules.py:266-279 / :215-230 (DSG :420-431, :465-479) for every layer at once.
    Returns a list of dicts {wq, lo, hi, codes}.
    """
    n = len(weights)
    descs = (N.WeightDesc * max(n, 1))()
    results = []
    device = None
    keep = []
    for i, w in enumerate(weights):
        _need(w, f"weight[{i}]")
        if w.dim() < 1:
            raise RuntimeError("ood_dfq_b200: weight must have at least one dimension")
        device = device or w.device
        if w.device != device:
            raise RuntimeError("ood_dfq_b200: all weights of one launch must live on the same device")
        wc = w.detach()
        if not (wc.is_contiguous() or (wc.dim() == 4 and wc.is_contiguous(memory_format=torch.channels_last))):
            wc = wc.contiguous()
        # dim 0 is the slowest axis in both dense formats, so an output row is one contiguous run either way
        keep.append(wc)
        rows = wc.shape[0]
        row_len = wc.numel() // rows if rows else 0
        wq = outs[i] if outs is not None else torch.empty_like(wc)
        lo = torch.empty(rows, dtype=torch.float32, device=device) if want_range else None
        hi = torch.empty(rows, dtype=torch.float32, device=device) if want_range else None
        cd = torch.empty_like(wc, dtype=torch.int8) if want_codes else None
        d = descs[i]
        d.w, d.wq, d.lo, d.hi, d.codes = wc.data_ptr(), wq.data_ptr(), _ptr(lo), _ptr(hi), _ptr(cd)
        d.rows, d.row_len, d.k = rows, row_len, int(ks[i])
        d.flags = N.SYMMETRIC if symmetric[i] else 0
        results.append({"wq": wq, "lo": lo, "hi": hi, "codes": cd})
    if n:
        rc = N.load().oodfq_weight_fq_multi(descs, n, _stream(device))
        N.check(rc, "weight_fq_multi")
    return results


# ----------------------------------------------------------------------------- a11 / a12
def _nchw_or_nhwc(x: torch.Tensor):
    """(dense tensor, N, C, HW, nhwc?) -- channels_last tensors are used as they are (C % 4 == 0)."""
    if x.dim() != 4:
        raise RuntimeError(f"ood_dfq_b200: expected an NCHW / channels_last 4-D tensor, got {x.dim()}-D")
    n, c, h, w = x.shape
    if (not x.is_contiguous()) and x.is_contiguous(memory_format=torch.channels_last) and c % 4 == 0:
        return x, n, c, h * w, True
    return x.contiguous(), n, c, h * w, False


def bn_stats_forward(x, shift=None, sums=None, fq=None):
    """Shifted per-channel sums of a BN input in one read (trainer_direct.py:388-393).

    Returns fp64 sums [2*C] (S1 then S2), or (sums, y) when ``fq=(k, lo, hi)`` asks for the fused
    fake-quantised output of the same read.
    """
    _need(x, "input")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    if shift is not None:
        _need(shift, "shift")
        if shift.numel() != c or not shift.is_contiguous():
            raise RuntimeError("ood_dfq_b200: shift must be a contiguous [C] tensor")
    if sums is None:
        sums = torch.empty(2 * c, dtype=torch.float64, device=x.device)
    elif sums.dtype != torch.float64 or sums.numel() != 2 * c or not sums.is_contiguous():
        raise RuntimeError("ood_dfq_b200: sums must be a contiguous float64 [2*C] tensor")
    _keep_until_flush(sums)                 # (inside ops.deferred_folds the sums are written by the flush)
    y = lo = hi = None
    k = 0
    if fq is not None:
        k, lo, hi = fq
        _need(lo, "fq range min")
        _need(hi, "fq range max")
        y = torch.empty_like(xc)
    name = "bn_*_stats_kernel + fused fake-quant (8 B/elem)" if fq is not None else \
        "bn_*_stats_kernel (BN-input statistics, 4 B/elem)"
    with _Timed(name, (8 if fq is not None else 4) * xc.numel()):
        rc = N.load().oodfq_bn_stats_forward(xc.data_ptr(), n, c, hw, _ptr(shift), sums.data_ptr(), _ptr(y),
                                             _ptr(lo), _ptr(hi), int(k), N.BN_NHWC if nhwc else 0,
                                             workspace(x.device).data_ptr(), _stream(x.device))
        N.check(rc, "bn_stats_forward")
    return (sums, y) if fq is not None else sums


def bn_stats_finalize(sums, shift, count: float):
    c = sums.numel() // 2
    mean = torch.empty(c, dtype=torch.float32, device=sums.device)
    var = torch.empty(c, dtype=torch.float32, device=sums.device)
    rc = N.load().oodfq_bn_stats_finalize(sums.data_ptr(), _ptr(shift), c, float(count), mean.data_ptr(),
                                          var.data_ptr(), _stream(sums.device))
    N.check(rc, "bn_stats_finalize")
    return mean, var


def bns_loss(sums, shift, run_mean, run_var, ch_off: Sequence[int], counts: Sequence[float]):
    """Packed BN-statistics loss over L layers (trainer_direct.py:473-486, distill_data.py:252-265).

    Returns (loss3, mean, var, gmean, gvar); loss3 = [total, mean term, var term].
    """
    L = len(counts)
    ctot = ch_off[-1]
    dev = sums.device
    loss3 = torch.empty(3, dtype=torch.float32, device=dev)
    mean, var, gmean, gvar = (torch.empty(ctot, dtype=torch.float32, device=dev) for _ in range(4))
    off = (C.c_int * (L + 1))(*ch_off)
    cnt = (C.c_double * L)(*counts)
    rc = N.load().oodfq_bns_loss(sums.data_ptr(), _ptr(shift), run_mean.data_ptr(), run_var.data_ptr(), off, cnt,
                                 L, loss3.data_ptr(), mean.data_ptr(), var.data_ptr(), gmean.data_ptr(),
                                 gvar.data_ptr(), _stream(dev))
    N.check(rc, "bns_loss")
    return loss3, mean, var, gmean, gvar


def bn_stats_backward(x, grad_in, mean, gmean, gvar, count: float, gscale=None, out=None):
    """grad_x = grad_in + g*(gmean_c/M + gvar_c*2(x-mean_c)/M); one fused pass."""
    _need(x, "input")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    gi = None
    if grad_in is not None:
        _need(grad_in, "grad_in")
        gi = grad_in.contiguous(memory_format=torch.channels_last) if nhwc else grad_in.contiguous()
    gx = out if out is not None else torch.empty_like(xc)
    with _Timed("bn_*_bwd_kernel (BNS-loss backward, 12 B/elem accumulating, 8 fresh)",
                (12 if gi is not None else 8) * xc.numel()):
        rc = N.load().oodfq_bn_stats_backward(xc.data_ptr(), _ptr(gi), gx.data_ptr(), n, c, hw, mean.data_ptr(),
                                              gmean.data_ptr(), gvar.data_ptr(), float(count), _ptr(gscale),
                                              N.BN_NHWC if nhwc else 0, _stream(x.device))
        N.check(rc, "bn_stats_backward")
    return gx


# ----------------------------------------------------------------------------- 8(f)-1: fused eval-mode BN
def _bn_ptrs(weight, bias, running_mean, running_var, c):
    for t, nme in ((running_mean, "running_mean"), (running_var, "running_var")):
        _need(t, nme)
    for t, nme in ((weight, "weight"), (bias, "bias"), (running_mean, "running_mean"), (running_var, "running_var")):
        if t is not None and (t.numel() != c or not t.is_contiguous() or t.dtype != torch.float32):
            raise RuntimeError(f"ood_dfq_b200: BN {nme} must be a contiguous fp32 [C] tensor")
    return _ptr(weight), _ptr(bias), running_mean.data_ptr(), running_var.data_ptr()


def bn_eval_forward(x, weight, bias, running_mean, running_var, eps, relu=False, fq=None, want_z=False, want_mask=False):
    """``[fakequant]([relu](BN_eval(x)))`` in one pass.  ``fq = (k, lo, hi)`` with a scalar range.

    NCHW-contiguous and channels_last inputs both run natively; the output keeps the input's memory format.
    ``want_mask`` (channels_last + ``relu``): also return the one-byte-per-four-channels ReLU mask that lets
    ``bn_eval_backward`` skip the read of ``x`` when no parameter gradients are wanted (None where unsupported).
    Returns ``y``, or a tuple ``(y[, z][, mask])`` in that order for the extras asked for.
    """
    _need(x, "input")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    y = torch.empty_like(xc)
    z = torch.empty_like(xc) if want_z else None
    mask = torch.empty(xc.numel() // 4, dtype=torch.uint8, device=x.device) if (want_mask and nhwc and relu) else None
    flags, k, lo, hi = (N.BN_RELU if relu else 0) | (N.BN_NHWC if nhwc else 0), 0, None, None
    if fq is not None:
        k, lo, hi = fq
        _need(lo, "fq range min")
        _need(hi, "fq range max")
        flags |= N.BN_QUANT
    name = "bn_*_fwd_kernel<relu,quant> (BN+ReLU+QuantAct forward, 8 B/elem)" if fq is not None else \
        "bn_*_fwd_kernel (eval BN forward, 8 B/elem)"
    with _Timed(name, 8 * xc.numel() + (xc.numel() // 4 if mask is not None else 0)):
        rc = N.load().oodfq_bn_eval_forward(xc.data_ptr(), y.data_ptr(), _ptr(z), n, c, hw, pw, pb, prm, prv,
                                            float(eps), flags, _ptr(lo), _ptr(hi), int(k), _ptr(mask), _stream(x.device))
        N.check(rc, "bn_eval_forward")
    if not (want_z or want_mask):
        return y
    return (y,) + ((z,) if want_z else ()) + ((mask,) if want_mask else ())


def bn_eval_backward(x, grad_y, weight, bias, running_mean, running_var, eps, relu=False, want_param_grads=True, mask=None):
    """Backward of ``bn_eval_forward`` (identity STE through the quantiser): (grad_x, dweight, dbias).
    ``mask``: the forward's ReLU mask; with it and without parameter gradients ``x`` is not read (may be None)."""
    _need(grad_y, "grad_output")
    by_mask = mask is not None and relu and not want_param_grads
    if x is None and not by_mask:
        raise RuntimeError("ood_dfq_b200: bn_eval_backward needs x unless the ReLU mask replaces it")
    if by_mask:
        _need(mask, "relu mask", torch.uint8)
        gc, n, c, hw, nhwc = _nchw_or_nhwc(grad_y)
        if not nhwc or mask.numel() != gc.numel() // 4 or not mask.is_contiguous():
            raise RuntimeError("ood_dfq_b200: the ReLU mask belongs to a channels_last tensor of the gradient's shape")
        xc, gy = None, gc
    else:
        _need(x, "input")
        xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
        gy = grad_y.contiguous(memory_format=torch.channels_last) if nhwc else grad_y.contiguous()
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    gx = torch.empty_like(gy)
    dwdb = _keep_until_flush(torch.empty(2 * c, dtype=torch.float32, device=gy.device) if want_param_grads else None)
    ws = workspace(gy.device).data_ptr() if want_param_grads else None
    reads_x = (relu or want_param_grads) and not by_mask   # otherwise grad_x = grad_y * a_c and x is never touched
    with _Timed("bn_*_bwdx_kernel (fused BN backward, 12 B/elem; 8 without ReLU mask and parameter grads)",
                int((12 if reads_x else (8.25 if by_mask else 8)) * gy.numel())):
        rc = N.load().oodfq_bn_eval_backward(_ptr(xc), gy.data_ptr(), gx.data_ptr(), n, c, hw, pw, pb, prm, prv,
                                             float(eps), (N.BN_RELU if relu else 0) | (N.BN_NHWC if nhwc else 0),
                                             _ptr(dwdb), ws, _ptr(mask) if by_mask else None, _stream(gy.device))
        N.check(rc, "bn_eval_backward")
    if not want_param_grads:
        return gx, None, None
    d = dwdb
    return gx, d[:c], d[c:]


def bn_eval_stats_forward(x, weight, bias, running_mean, running_var, eps, shift, sums, relu=False, fq=None):
    """``bn_stats_forward(x, shift, sums=sums)`` and ``bn_eval_forward(x, ...)`` from ONE read of a channels_last ``x``:
    returns ``y`` (bit-identical to ``bn_eval_forward``) and fills ``sums`` (fp64 ``[2*C]``, shifted by ``shift``)."""
    _need(x, "input")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    if not nhwc:
        raise RuntimeError("ood_dfq_b200: bn_eval_stats_forward takes channels_last tensors (C % 4 == 0) only")
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    if sums.dtype != torch.float64 or sums.numel() != 2 * c or not sums.is_contiguous():
        raise RuntimeError("ood_dfq_b200: sums must be a contiguous float64 [2*C] tensor")
    _keep_until_flush(sums)
    y = torch.empty_like(xc)
    flags, k, lo, hi = (N.BN_RELU if relu else 0) | N.BN_NHWC, 0, None, None
    if fq is not None:
        k, lo, hi = fq
        flags |= N.BN_QUANT
    with _Timed("bn_*_stats_kernel + fused BN forward (8 B/elem)", 8 * xc.numel()):
        rc = N.load().oodfq_bn_eval_stats_forward(xc.data_ptr(), y.data_ptr(), n, c, hw, pw, pb, prm, prv, float(eps), flags,
                                                  _ptr(lo), _ptr(hi), int(k), _ptr(shift), sums.data_ptr(),
                                                  workspace(x.device).data_ptr(), _stream(x.device))
        N.check(rc, "bn_eval_stats_forward")
    return y


def bn_eval_tap_backward(x, grad_y, weight, bias, running_mean, running_var, eps, mean, gmean, gvar, count: float,
                         relu=False, gscale=None):
    """``bn_eval_backward`` (no parameter gradients) followed by ``bn_stats_backward`` accumulating into its result, as
    ONE pass: the backward of a fused BatchNorm whose input is tapped by the BN-statistics loss.  channels_last only."""
    _need(x, "input")
    _need(grad_y, "grad_output")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    if not nhwc:
        raise RuntimeError("ood_dfq_b200: bn_eval_tap_backward takes channels_last tensors (C % 4 == 0) only")
    gy = grad_y.contiguous(memory_format=torch.channels_last)
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    gx = torch.empty_like(xc)
    with _Timed("bn_*_bwdx_tap_kernel (fused BN backward + BNS-loss backward, 12 B/elem)", 12 * xc.numel()):
        rc = N.load().oodfq_bn_eval_tap_backward(xc.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, c, hw, pw, pb, prm, prv,
                                                 float(eps), (N.BN_RELU if relu else 0) | N.BN_NHWC, mean.data_ptr(),
                                                 gmean.data_ptr(), gvar.data_ptr(), float(count), _ptr(gscale),
                                                 _stream(x.device))
        N.check(rc, "bn_eval_tap_backward")
    return gx


# ----------------------------------------------------------------------------- 8(f)-2: feature-alignment reduction
def channel_energy_forward(x):
    """``x.pow(2).mean([2, 3])`` of an NCHW / channels_last tensor in one read (trainer_direct.py:382-383)."""
    _need(x, "input")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    e = torch.empty((n, c), dtype=torch.float32, device=x.device)
    scratch = None
    if nhwc:
        scratch = torch.empty(int(N.load().oodfq_channel_energy_scratch_floats(n, c)), dtype=torch.float32,
                              device=x.device)
    with _Timed("energy_*_kernel (feature-alignment mean of squares, 4 B/elem)", 4 * xc.numel()):
        rc = N.load().oodfq_channel_energy_forward(xc.data_ptr(), e.data_ptr(), n, c, hw, N.BN_NHWC if nhwc else 0,
                                                   _ptr(scratch), _stream(x.device))
        N.check(rc, "channel_energy_forward")
    return e


def channel_energy_backward(x, grad_e):
    """``grad_x = grad_e[n,c] * 2/HW * x``; keeps x's memory format."""
    _need(x, "input")
    _need(grad_e, "grad_e")
    xc, n, c, hw, nhwc = _nchw_or_nhwc(x)
    ge = grad_e.contiguous()
    gx = torch.empty_like(xc)
    with _Timed("energy_*_bwd (feature-alignment backward, 8 B/elem)", 8 * xc.numel()):
        rc = N.load().oodfq_channel_energy_backward(xc.data_ptr(), ge.data_ptr(), gx.data_ptr(), n, c, hw,
                                                    N.BN_NHWC if nhwc else 0, _stream(x.device))
        N.check(rc, "channel_energy_backward")
    return gx


# ----------------------------------------------------------------------------- global average pool (final_pool)
def global_avgpool_supported(x) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0
            and x.shape[1] % 4 == 0 and x.shape[2] * x.shape[3] > 1
            and x.is_contiguous(memory_format=torch.channels_last))


def global_avgpool_forward(x):
    """``avg_pool2d(x, (H, W))`` of a channels_last tensor -> ``[N, C, 1, 1]``: fp32 running sum in window order and
    one division, i.e. the bits ATen's kernel produces."""
    _need(x, "input")
    if not global_avgpool_supported(x):
        raise RuntimeError("ood_dfq_b200: the global average pool needs a channels_last fp32 tensor with C % 4 == 0")
    n, c, h, w = x.shape
    y = torch.empty((n, c, 1, 1), dtype=torch.float32, device=x.device)
    with _Timed("gap_nhwc_fwd_kernel (global average pool, 4 B/elem)", 4 * x.numel()):
        rc = N.load().oodfq_global_avgpool_forward(x.data_ptr(), y.data_ptr(), n, c, h * w, N.BN_NHWC, _stream(x.device))
        N.check(rc, "global_avgpool_forward")
    return y


def global_avgpool_backward(grad_y, in_shape):
    """``grad_x[n,c,h,w] = grad_y[n,c] / (H*W)`` as a channels_last tensor of ``in_shape``."""
    _need(grad_y, "grad_output")
    n, c, h, w = in_shape
    gy = grad_y.reshape(n, c).contiguous()
    gx = torch.empty((n, c, h, w), dtype=torch.float32, device=grad_y.device, memory_format=torch.channels_last)
    with _Timed("gap_nhwc_bwd_kernel (global average pool backward, 4 B/elem)", 4 * gx.numel()):
        rc = N.load().oodfq_global_avgpool_backward(gy.data_ptr(), gx.data_ptr(), n, c, h * w, N.BN_NHWC,
                                                    _stream(grad_y.device))
        N.check(rc, "global_avgpool_backward")
    return gx


# ----------------------------------------------------------------------------- 8(f)-2: feature-alignment loss
def _fa_tables(es, et):
    if len(es) != len(et) or not es:
        raise RuntimeError("ood_dfq_b200: the feature-alignment loss needs as many student as teacher maps (>= 1)")
    if len(es) > int(N.load().oodfq_fa_loss_max_layers()):
        raise RuntimeError(f"ood_dfq_b200: at most {int(N.load().oodfq_fa_loss_max_layers())} residual units per call")
    n = es[0].shape[0]
    keep = []
    for i, (a, b) in enumerate(zip(es, et)):
        _need(a, f"student energy[{i}]")
        _need(b, f"teacher energy[{i}]")
        if a.dim() != 2 or a.shape != b.shape or a.shape[0] != n:
            raise RuntimeError("ood_dfq_b200: energies must be [N, C_l] tensors, pairwise of one shape")
        keep.append((a.contiguous(), b.contiguous()))
    L = len(keep)
    ps = (C.c_void_p * L)(*[a.data_ptr() for a, _ in keep])
    pt = (C.c_void_p * L)(*[b.data_ptr() for _, b in keep])
    ch = (C.c_int * L)(*[a.shape[1] for a, _ in keep])
    return keep, ps, pt, ch, L, n


def fa_loss_forward(es, et, lam: float):
    """``lam * sum_l mean((F.normalize(es[l]) - F.normalize(et[l]))**2)`` (trainer_direct.py:325-330, :382-383) of all
    residual units in one launch.  ``es`` / ``et``: lists of ``[N, C_l]`` energies.  Returns a 1-element tensor."""
    keep, ps, pt, ch, L, n = _fa_tables(es, et)
    dev = keep[0][0].device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    scratch = torch.empty(L * n, dtype=torch.float64, device=dev)
    rc = N.load().oodfq_fa_loss_forward(ps, pt, ch, L, n, float(lam), loss.data_ptr(), scratch.data_ptr(),
                                        workspace(dev).data_ptr(), _stream(dev))
    N.check(rc, "fa_loss_forward")
    return loss


def fa_loss_backward(es, et, lam: float, grad_loss, want_student=True, want_teacher=True):
    """Gradients of ``fa_loss_forward`` w.r.t. every energy: ``(list for es, list for et)``; a side that is not
    wanted comes back as a list of None."""
    keep, ps, pt, ch, L, n = _fa_tables(es, et)
    dev = keep[0][0].device
    if grad_loss is not None:
        _need(grad_loss, "grad_loss")
        grad_loss = grad_loss.reshape(-1)[:1].contiguous()
    gs = [torch.empty_like(a) for a, _ in keep] if want_student else [None] * L
    gt = [torch.empty_like(b) for _, b in keep] if want_teacher else [None] * L
    pgs = (C.c_void_p * L)(*[_ptr(g) for g in gs]) if want_student else None
    pgt = (C.c_void_p * L)(*[_ptr(g) for g in gt]) if want_teacher else None
    if want_student or want_teacher:
        rc = N.load().oodfq_fa_loss_backward(ps, pt, ch, L, n, float(lam), _ptr(grad_loss), pgs, pgt, _stream(dev))
        N.check(rc, "fa_loss_backward")
    return gs, gt


# ----------------------------------------------------------------------------- stem: BN -> ReLU -> [QuantAct] -> MaxPool(3,2,1)
def bn_pool_supported(x) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0 and x.shape[1] <= 1024
            and x.is_contiguous(memory_format=torch.channels_last))


def bn_pool_forward(x, weight, bias, running_mean, running_var, eps, fq=None, want_xhat=True, register_kernel=False):
    """``max_pool2d(fakequant(relu(BN_eval(x))), 3, 2, 1)`` without ever writing the full-resolution tensor.

    channels_last input only.  Returns (out, idx, xhat): idx is the one-byte argmax code the backward needs,
    xhat the normalised input at the argmax (None unless ``want_xhat``; only BN parameter gradients use it).
    ``register_kernel`` forces the register-staged kernel (the fallback for rows too wide for the TMA ring;
    tests and comparisons).
    """
    _need(x, "input")
    if not bn_pool_supported(x):
        raise RuntimeError("ood_dfq_b200: the fused stem needs a channels_last fp32 tensor with C % 4 == 0")
    n, c, h, w = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    idx = torch.empty((n, c, ho, wo), dtype=torch.uint8, device=x.device, memory_format=torch.channels_last)
    xhat = torch.empty_like(out) if want_xhat else None
    flags, k, lo, hi = N.BN_RELU | N.BN_NHWC, 0, None, None
    if fq is not None:
        k, lo, hi = fq
        flags |= N.BN_QUANT
    if register_kernel:
        flags |= N.BN_POOL_REGISTER
    with _Timed("bn_pool_fwd_kernel (BN+ReLU+QuantAct+MaxPool stem forward, 4 B/elem in + 1/4 size outputs)",
                4 * x.numel() + (9 if want_xhat else 5) * out.numel()):
        rc = N.load().oodfq_bn_pool_forward(x.data_ptr(), out.data_ptr(), idx.data_ptr(), _ptr(xhat), n, c, h, w,
                                            pw, pb, prm, prv, float(eps), flags, _ptr(lo), _ptr(hi), int(k),
                                            _stream(x.device))
        N.check(rc, "bn_pool_forward")
    return out, idx, xhat


def bn_pool_backward(grad_out, idx, xhat, in_shape, weight, bias, running_mean, running_var, eps,
                     want_param_grads=True, grad_out2=None):
    """Backward of ``bn_pool_forward``: (grad_x channels_last, dweight, dbias).  The forward input is not needed.
    ``grad_out2``: a second gradient w.r.t. the output (it fed two consumers), summed inside the kernel."""
    _need(grad_out, "grad_output")
    n, c, h, w = in_shape
    go = grad_out.contiguous(memory_format=torch.channels_last)
    go2 = None
    if grad_out2 is not None:
        _need(grad_out2, "grad_output (second)")
        go2 = grad_out2.contiguous(memory_format=torch.channels_last)
    pw, pb, prm, prv = _bn_ptrs(weight, bias, running_mean, running_var, c)
    gx = torch.empty((n, c, h, w), dtype=torch.float32, device=grad_out.device, memory_format=torch.channels_last)
    need = want_param_grads and xhat is not None
    dwdb = _keep_until_flush(torch.empty(2 * c, dtype=torch.float32, device=grad_out.device) if need else None)
    ws = workspace(grad_out.device).data_ptr() if need else None
    with _Timed("bn_pool_bwd_kernel (stem backward, 4 B/elem out + 1/4 size inputs)",
                4 * gx.numel() + ((9 if need else 5) + (4 if go2 is not None else 0)) * go.numel()):
        rc = N.load().oodfq_bn_pool_backward(go.data_ptr(), _ptr(go2), idx.data_ptr(), _ptr(xhat) if need else None, gx.data_ptr(),
                                             n, c, h, w, pw, pb, prm, prv, float(eps), _ptr(dwdb), ws,
                                             _stream(grad_out.device))
        N.check(rc, "bn_pool_backward")
    if not need:
        return gx, None, None
    d = dwdb
    return gx, d[:c], d[c:]


# ----------------------------------------------------------------------------- residual-unit tail
def res_tail_supported(x1, r) -> bool:
    """channels_last fp32 CUDA tensors of one shape with C % 4 == 0 and C <= 1024."""
    return (isinstance(x1, torch.Tensor) and isinstance(r, torch.Tensor) and x1.is_cuda and r.is_cuda
            and x1.dtype == torch.float32 and r.dtype == torch.float32 and x1.dim() == 4 and x1.shape == r.shape
            and x1.shape[1] % 4 == 0 and x1.shape[1] <= 1024 and x1.numel() > 0
            and x1.is_contiguous(memory_format=torch.channels_last)
            and r.is_contiguous(memory_format=torch.channels_last))


def _tail_bn(bn, c):
    """(w, b, rm, rv, eps) pointers of an eval-mode BatchNorm given as a tuple, or NULLs for 'no BatchNorm'."""
    if bn is None:
        return None, None, None, None, 0.0
    weight, bias, running_mean, running_var, eps = bn
    return _bn_ptrs(weight, bias, running_mean, running_var, c) + (float(eps),)


def res_tail_forward(x1, r, bn1, bn2=None, fq=None, want_energy=False, want_mask=False):
    """``y = [fakequant](relu(BN1(x1) + id))`` with ``id = r`` or ``BN2(r)``, and optionally the per-(image,
    channel) mean of squares of ``BN1(x1)`` (the feature-alignment tap of the body output) from the same read.

    ``bn1`` / ``bn2``: ``(weight, bias, running_mean, running_var, eps)``; ``fq = (k, lo, hi)``.
    Returns ``(y, energy or None)``; with ``want_mask`` ``(y, energy or None, mask)`` where ``mask`` is one byte per
    four channels saying where the ReLU is open -- handed to ``res_tail_backward`` it spares that pass the read of
    ``r`` (and of ``x1`` when neither energy nor parameter gradients are wanted).
    """
    _need(x1, "body output")
    _need(r, "identity")
    if not res_tail_supported(x1, r):
        raise RuntimeError("ood_dfq_b200: the fused residual tail needs two channels_last fp32 tensors of one shape "
                           "with C % 4 == 0 and C <= 1024")
    n, c, h, w = x1.shape
    p1, p2 = _tail_bn(bn1, c), _tail_bn(bn2, c)
    y = torch.empty_like(x1)
    energy = scratch = None
    if want_energy:
        energy = torch.empty((n, c), dtype=torch.float32, device=x1.device)
        scratch = torch.empty(int(N.load().oodfq_res_tail_scratch_floats(n, c)), dtype=torch.float32, device=x1.device)
    mask = torch.empty(x1.numel() // 4, dtype=torch.uint8, device=x1.device) if want_mask else None
    flags, k, lo, hi = N.BN_NHWC, 0, None, None
    if fq is not None:
        k, lo, hi = fq
        flags |= N.BN_QUANT
    with _Timed("res_tail_fwd_kernel (BN + residual add + ReLU + QuantAct [+ energy], 12 B/elem)",
                12 * x1.numel() + (x1.numel() // 4 if want_mask else 0)):
        rc = N.load().oodfq_res_tail_forward(x1.data_ptr(), r.data_ptr(), y.data_ptr(), _ptr(energy), _ptr(scratch),
                                             _ptr(mask), n, c, h * w, *p1, *p2, flags, _ptr(lo), _ptr(hi), int(k),
                                             _stream(x1.device))
        N.check(rc, "res_tail_forward")
    return (y, energy, mask) if want_mask else (y, energy)


def res_tail_backward(grad_y, grad_energy, x1, r, bn1, bn2=None, want_param_grads=True, grad_y2=None, mask=None):
    """Backward of ``res_tail_forward``: ``(grad_x1, grad_r, dW1, dB1, dW2, dB2)`` (parameter gradients None
    unless wanted; dW2 / dB2 None without ``bn2``).  ``grad_y2``: a second gradient w.r.t. ``y`` (the output fed
    two consumers), added to ``grad_y`` inside the kernel.  ``mask``: the forward's ReLU mask; with it ``r`` may be
    None unless ``bn2``'s parameter gradients are wanted, ``x1`` may be None without energy / parameter gradients."""
    _need(grad_y, "grad_output")
    n, c, h, w = grad_y.shape
    gy = grad_y.contiguous(memory_format=torch.channels_last)
    gy2 = None
    if grad_y2 is not None:
        _need(grad_y2, "grad_output (second)")
        gy2 = grad_y2.contiguous(memory_format=torch.channels_last)
    ge = None
    if grad_energy is not None:
        _need(grad_energy, "grad_energy")
        ge = grad_energy.contiguous()
    if mask is not None:
        _need(mask, "relu mask", torch.uint8)
        if mask.numel() != gy.numel() // 4 or not mask.is_contiguous():
            raise RuntimeError("ood_dfq_b200: the ReLU mask must be the forward's contiguous [numel / 4] byte tensor")
    reads_x1 = mask is None or ge is not None or want_param_grads
    reads_r = mask is None or (bn2 is not None and want_param_grads)
    if (reads_x1 and x1 is None) or (reads_r and r is None):
        raise RuntimeError("ood_dfq_b200: res_tail_backward needs x1 / r for this combination (see the docstring)")
    p1, p2 = _tail_bn(bn1, c), _tail_bn(bn2, c)
    gx1, gr = torch.empty_like(gy), torch.empty_like(gy)
    ct = c * (2 if bn2 is not None else 1)
    dwdb = _keep_until_flush(torch.empty(2 * ct, dtype=torch.float32, device=gy.device) if want_param_grads else None)
    ws = workspace(gy.device).data_ptr() if want_param_grads else None
    per_elem = 12 + (4 if gy2 is not None else 0) + (4 if reads_x1 else 0) + (4 if reads_r else 0) + (0.25 if mask is not None else 0)
    with _Timed("res_tail_bwd_kernel (ReLU mask + energy gradient + BN backward(s), 16.25 - 24 B/elem)",
                int(per_elem * gy.numel())):
        rc = N.load().oodfq_res_tail_backward(gy.data_ptr(), _ptr(gy2), _ptr(ge), _ptr(x1) if reads_x1 else None,
                                              _ptr(r) if reads_r else None, _ptr(mask), gx1.data_ptr(),
                                              gr.data_ptr(), n, c, h * w, *p1, *p2, N.BN_NHWC, _ptr(dwdb), ws,
                                              _stream(gy.device))
        N.check(rc, "res_tail_backward")
    if not want_param_grads:
        return gx1, gr, None, None, None, None
    d = dwdb
    if bn2 is None:
        return gx1, gr, d[:c], d[c:], None, None
    return gx1, gr, d[:c], d[2 * c:3 * c], d[c:2 * c], d[3 * c:]


# ----------------------------------------------------------------------------- stem convolution input re-layout
def s2d_stem_supported(x, pad) -> bool:
    return (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0
            and (x.shape[2] + 2 * pad) % 2 == 0 and (x.shape[3] + 2 * pad) % 2 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


def s2d_stem_forward(x, pad, cpad=None):
    """channels_last ``[N,C,H,W]`` -> channels_last ``[N,CP,(H+2p)/2,(W+2p)/2]`` with channel order (s, t, c):
    ``xs[n,(s,t,c),i,j] = x[n,c,2i+s-p,2j+t-p]``, zero outside the image.  ``CP = cpad`` (default ``4C``): a larger
    multiple of 4 appends zero channels (cuDNN takes 16-channel tensors as they are, 12-channel ones it converts)."""
    _need(x, "input")
    if not s2d_stem_supported(x, pad):
        raise RuntimeError("ood_dfq_b200: space-to-depth needs a channels_last fp32 CUDA tensor with even H+2p, W+2p")
    n, c, h, w = x.shape
    cp = 4 * c if cpad is None else int(cpad)
    if cp < 4 * c or cp % 4:
        raise RuntimeError("ood_dfq_b200: cpad must be a multiple of 4 that is >= 4*C")
    xs = torch.empty((n, cp, (h + 2 * pad) // 2, (w + 2 * pad) // 2), dtype=torch.float32, device=x.device,
                     memory_format=torch.channels_last)
    with _Timed("s2d_stem_kernel (stem input re-layout, 8 B/elem)", 4 * (x.numel() + xs.numel())):
        rc = N.load().oodfq_s2d_stem_forward(x.data_ptr(), xs.data_ptr(), n, h, w, c, int(pad), cp, _stream(x.device))
        N.check(rc, "s2d_stem_forward")
    return xs


def s2d_stem_backward(grad_xs, in_shape, pad):
    """Gradient of ``s2d_stem_forward`` w.r.t. its input (channels_last ``in_shape``); the channel padding, if any, is
    read off ``grad_xs``."""
    _need(grad_xs, "grad_output")
    n, c, h, w = in_shape
    g = grad_xs.contiguous(memory_format=torch.channels_last)
    cp = g.shape[1]
    if cp < 4 * c or cp % 4:
        raise RuntimeError("ood_dfq_b200: grad_output must have 4*C channels (or more, as a multiple of 4)")
    gx = torch.empty((n, c, h, w), dtype=torch.float32, device=g.device, memory_format=torch.channels_last)
    with _Timed("s2d_stem_kernel (stem input re-layout, 8 B/elem)", 4 * (gx.numel() + g.numel())):
        rc = N.load().oodfq_s2d_stem_backward(g.data_ptr(), gx.data_ptr(), n, h, w, c, int(pad), cp, _stream(g.device))
        N.check(rc, "s2d_stem_backward")
    return gx


# ----------------------------------------------------------------------------- QuantAct_MSE range search
def act_mse_search(x, k, x_min, x_max, beta, beta_t, cur_min=None, cur_max=None, steps=80, step=0.01, p=2.4,
                   debug=False):
    """The clip-ratio search of ``QuantAct_MSE.forward`` (quant_modules.py:160-178) on the device: data min/max,
    all ``steps`` candidates scored in one pass over ``x``, first strict minimum kept, plain EMA into
    ``x_min / x_max / beta_t`` in place.  No host synchronisation.  ``debug``: also return (scores, chosen)."""
    _need(x, "input")
    for t, nme in ((x_min, "x_min"), (x_max, "x_max"), (beta, "beta"), (beta_t, "beta_t")):
        _need(t, nme)
        if t.numel() != 1 or not t.is_contiguous():
            raise RuntimeError(f"ood_dfq_b200: {nme} must be a contiguous 1-element buffer")
    xd = _dense(x)
    mm = minmax(xd)
    scratch = torch.empty(int(N.load().oodfq_act_mse_scratch_doubles(int(steps))), dtype=torch.float64, device=x.device)
    scores = torch.empty(int(steps), dtype=torch.float32, device=x.device) if debug else None
    chosen = torch.empty(1, dtype=torch.int32, device=x.device) if debug else None
    rc = N.load().oodfq_act_mse_search(xd.data_ptr(), xd.numel(), mm.data_ptr(), int(k), int(steps), float(step),
                                       float(p), x_min.data_ptr(), x_max.data_ptr(), beta.data_ptr(),
                                       beta_t.data_ptr(), _ptr(cur_min), _ptr(cur_max), scratch.data_ptr(),
                                       _ptr(scores), _ptr(chosen), _stream(x.device))
    N.check(rc, "act_mse_search")
    return (scores, chosen) if debug else None


# ----------------------------------------------------------------------------- batch assembly (crop / resize / flip)
def _aug_source(images):
    """(flags, m, c, h, w) of an image set that is NCHW- or channels_last-contiguous."""
    if images.dim() != 4:
        raise RuntimeError("ood_dfq_b200: the image set must be a 4-D [M,C,H,W] tensor")
    m, c, h, w = images.shape
    if images.is_contiguous():
        return 0, m, c, h, w
    if images.is_contiguous(memory_format=torch.channels_last):
        return N.AUG_SRC_NHWC, m, c, h, w
    raise RuntimeError("ood_dfq_b200: the image set must be NCHW- or channels_last-contiguous")


def _aug_draws(index, boxes, flips):
    _need(index, "index", torch.int64)
    _need(boxes, "boxes", torch.int32)
    _need(flips, "flips", torch.uint8)
    n = index.numel()
    if index.dim() != 1 or boxes.shape != (n, 4) or flips.shape != (n,):
        raise RuntimeError(f"ood_dfq_b200: index [N], boxes [N,4], flips [N] expected, got {tuple(index.shape)}, "
                           f"{tuple(boxes.shape)}, {tuple(flips.shape)}")
    return n, index.contiguous(), boxes.contiguous(), flips.contiguous()


def crop_resize_flip(images, index, boxes, flips, size, channels=None, channels_last=True, out=None):
    """One per-rank batch straight from the device-resident image set: for sample ``n`` the crop ``boxes[n] = (top,
    left, h, w)`` of ``images[index[n]]`` is resized (bilinear, ``align_corners=False``) to ``size``, a one-channel
    image is repeated to ``channels`` (default 3) and the result mirrored when ``flips[n]`` -- the per-sample
    torchvision pipeline of ``direct_dataset`` (main_direct.py:158-169, :200-204) as one kernel.

    ``images`` ``[M,C,H,W]`` fp32 (NCHW- or channels_last-contiguous), ``index`` int64 ``[N]``, ``boxes`` int32
    ``[N,4]``, ``flips`` uint8 ``[N]``, all on the same CUDA device.  Returns ``[N, channels, *size]`` (channels_last
    by default).
    """
    _need(images, "image set")
    src_flag, m, c_in, h, w = _aug_source(images)
    n, index, boxes, flips = _aug_draws(index, boxes, flips)
    c_out = int(channels) if channels is not None else (3 if c_in == 1 else c_in)
    oh, ow = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    given = out is not None
    if out is None:
        out = torch.empty((n, c_out, oh, ow), dtype=torch.float32, device=images.device, memory_format=fmt)
    else:
        _need(out, "out")
        if out.shape != (n, c_out, oh, ow) or not out.is_contiguous(memory_format=fmt):
            raise RuntimeError("ood_dfq_b200: `out` must be a dense [N, channels, *size] tensor in the requested memory format")
    if n == 0:
        return out
    crop_bytes = 4 * c_in * h * w * n                         # upper bound of the read (the box is <= the image)
    with _Timed("crop_resize_flip_kernel (batch assembly, <= 4 B/elem in + 4 B/elem out)", crop_bytes + 4 * out.numel()):
        rc = N.load().oodfq_crop_resize_flip(images.data_ptr(), m, c_in, h, w, index.data_ptr(), boxes.data_ptr(),
                                             flips.data_ptr(), out.data_ptr(), n, c_out, oh, ow,
                                             (N.BN_NHWC if channels_last else 0) | src_flag, _stream(images.device))
        N.check(rc, "crop_resize_flip")
    return _written(out) if given else out


def crop_resize_flip_backward(grad_out, like, index, boxes, flips, accumulate_into=None):
    """Gradient of ``crop_resize_flip`` w.r.t. the image set: ``grad_out`` ``[N, C_out, OH, OW]`` (NCHW- or
    channels_last-contiguous) scattered onto a tensor shaped and laid out like ``like`` (the forward's ``images``)
    with the forward's tap weights.  ``accumulate_into``: add into this existing gradient instead of a zero tensor."""
    _need(grad_out, "grad_output")
    _need(like, "image set")
    src_flag, m, c_in, h, w = _aug_source(like)
    n, index, boxes, flips = _aug_draws(index, boxes, flips)
    if grad_out.dim() != 4 or grad_out.shape[0] != n:
        raise RuntimeError("ood_dfq_b200: grad_output must be [N, C_out, OH, OW] with one sample per index entry")
    if grad_out.is_contiguous(memory_format=torch.channels_last):
        out_flag = N.BN_NHWC
    else:
        grad_out, out_flag = grad_out.contiguous(), 0
    _, c_out, oh, ow = grad_out.shape
    if accumulate_into is None:
        grad_images = torch.zeros_like(like)                  # preserves the layout of `like`
    else:
        grad_images = accumulate_into
        if grad_images.shape != like.shape or grad_images.stride() != like.stride():
            raise RuntimeError("ood_dfq_b200: `accumulate_into` must have the shape and layout of the image set")
    if n == 0:
        return grad_images
    with _Timed("crop_resize_flip_bwd_kernel (gradient scatter, 4 B/elem in + atomics)", 4 * grad_out.numel() + 8 * c_in * h * w * n):
        rc = N.load().oodfq_crop_resize_flip_backward(grad_out.data_ptr(), grad_images.data_ptr(), m, c_in, h, w,
                                                      index.data_ptr(), boxes.data_ptr(), flips.data_ptr(), n, c_out, oh,
                                                      ow, out_flag | src_flag, _stream(like.device))
        N.check(rc, "crop_resize_flip_backward")
    return _written(grad_images) if accumulate_into is not None else grad_images
