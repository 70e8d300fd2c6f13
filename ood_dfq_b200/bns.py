"""BN-statistics matching loss on B200: fused statistics, loss and backward kernels.

Reference behaviour being replaced (paths relative to the reference tree):

* the forward hook on every BatchNorm module that takes ``mean([0,2,3])`` and
  ``var([0,2,3], unbiased=False)`` of the module INPUT -- trainer_direct.py:388-397
  (registered :418-423) and data_generate/distill_data.py:69-78 (registered :156-158);
* the loss built from those lists -- trainer_direct.py:473-486
  (``sum_l [MSE(mean_l, rm_l) + MSE(var_l, rv_l)] / L``) and distill_data.py:252-265
  (mean term / L + var term / L: the same number, reported as two parts);
* the autograd tape behind both (several more full passes per layer in backward).

Two levels of API:

``bn_channel_stats(x, shift=None)``
    differentiable ``(mean, var)`` of one NCHW tensor: one read forward, one fused
    read+write backward.  Drop it into the reference's hook in place of the two ATen
    reductions.

``BNStatLoss(model)``
    hooks every BN module, accumulates the shifted sums of all layers into one packed
    buffer (optionally all-reduced over the data-parallel group so the loss is the
    GLOBAL-batch loss), evaluates loss and per-channel gradients in one tiny kernel, and
    adds the loss gradient into the gradient that flows back through each BN input in
    one fused kernel per layer.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops


class _ChannelStats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, shift):
        n, c, h, w = x.shape
        count = float(n * h * w)
        xc = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous()
        if shift is None:
            # any sample of the channel is a good pivot for the one-pass variance
            shift = xc[0, :, 0, 0].contiguous()
        sums = ops.bn_stats_forward(xc, shift)
        mean, var = ops.bn_stats_finalize(sums, shift, count)
        ctx.save_for_backward(xc, mean)
        ctx.count = count
        ctx.set_materialize_grads(False)
        return mean, var

    @staticmethod
    def backward(ctx, gmean, gvar):
        xc, mean = ctx.saved_tensors
        if gmean is None and gvar is None:
            return None, None
        gmean = torch.zeros_like(mean) if gmean is None else gmean.contiguous()
        gvar = torch.zeros_like(mean) if gvar is None else gvar.contiguous()
        return ops.bn_stats_backward(xc, None, mean, gmean, gvar, ctx.count), None


def bn_channel_stats(x: torch.Tensor, shift: Optional[torch.Tensor] = None):
    """Per-channel batch mean and biased variance of an NCHW tensor, differentiable.

    Equivalent to ``x.mean([0,2,3]), x.var([0,2,3], unbiased=False)``
    (trainer_direct.py:390-392).  ``shift`` ([C], e.g. the BN running mean) is only a
    numerical pivot for the one-pass variance; the result does not depend on it.
    """
    return _ChannelStats.apply(x, shift)


class _Pass:
    """State of one hooked forward pass (several may be alive before their backward runs)."""

    __slots__ = ("sums", "counts", "fired", "tokens", "world", "loss3", "mean", "var", "gmean", "gvar")

    def __init__(self, n_layers, ctot, device):
        self.sums = torch.empty(2 * ctot, dtype=torch.float64, device=device)
        self.counts = [0.0] * n_layers
        self.fired = [0] * n_layers
        self.tokens = [None] * n_layers
        self.world = 1
        self.loss3 = self.mean = self.var = self.gmean = self.gvar = None


class _Tap(torch.autograd.Function):
    """Identity on a BN input that records its channel sums; backward adds the loss gradient."""

    @staticmethod
    def forward(ctx, x, mgr, idx, run):
        n, c, h, w = x.shape
        lay = mgr._layers[idx]
        if c != lay.channels:
            raise RuntimeError(f"BNStatLoss: layer {idx} saw {c} channels, expected {lay.channels}")
        xc = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous()
        ops.bn_stats_forward(xc, lay.module.running_mean, sums=run.sums[2 * lay.offset: 2 * (lay.offset + c)])
        run.counts[idx] = float(n * h * w)
        run.fired[idx] += 1
        ctx.save_for_backward(xc)
        ctx.lay, ctx.idx, ctx.run = lay, idx, run
        ctx.set_materialize_grads(False)
        token = torch.empty((), dtype=torch.float32, device=x.device)
        return x, token

    @staticmethod
    def backward(ctx, grad_x, grad_token):
        if grad_token is None:
            return grad_x, None, None, None
        (xc,) = ctx.saved_tensors
        lay, run = ctx.lay, ctx.run
        sl = slice(lay.offset, lay.offset + lay.channels)
        g = ops.bn_stats_backward(xc, grad_x, run.mean[sl], run.gmean[sl], run.gvar[sl],
                                  run.counts[ctx.idx] * run.world, gscale=grad_token.reshape(1).contiguous())
        return g, None, None, None


class _TapFusedBN(torch.autograd.Function):
    """``_Tap`` and the fused eval-mode BatchNorm behind it (``fusion.FusedEvalBN``) as one autograd node: forward is
    the statistics kernel plus the fused BN(+ReLU+QuantAct) kernel as before, but the backward -- BN backward, then the
    loss gradient added into its result -- is ONE pass over x and grad_y (12 B/elem instead of 8-12 + 12) when no
    BatchNorm parameter gradients are wanted (the hooked network of the distillation loop is frozen)."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, relu, qact, mgr, idx, run):
        n, c, h, w = x.shape
        lay = mgr._layers[idx]
        if c != lay.channels:
            raise RuntimeError(f"BNStatLoss: layer {idx} saw {c} channels, expected {lay.channels}")
        xc = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous()
        sums = run.sums[2 * lay.offset: 2 * (lay.offset + c)]
        fq = (qact.activation_bit, qact.x_min, qact.x_max) if qact is not None else None
        if xc.is_contiguous(memory_format=torch.channels_last) and not xc.is_contiguous() and c % 4 == 0:
            # statistics and BatchNorm from ONE read of x (8 instead of 4 + 8 B/elem)
            y = ops.bn_eval_stats_forward(xc, weight, bias, bn.running_mean, bn.running_var, bn.eps,
                                          lay.module.running_mean, sums, relu=relu, fq=fq)
        else:
            ops.bn_stats_forward(xc, lay.module.running_mean, sums=sums)
            y = ops.bn_eval_forward(xc, weight, bias, bn.running_mean, bn.running_var, bn.eps, relu=relu, fq=fq)
        run.counts[idx] = float(n * h * w)
        run.fired[idx] += 1
        ctx.save_for_backward(xc, weight, bias)
        ctx.bn, ctx.relu, ctx.lay, ctx.idx, ctx.run = bn, relu, lay, idx, run
        ctx.set_materialize_grads(False)
        token = torch.empty((), dtype=torch.float32, device=x.device)
        return y, token

    @staticmethod
    def backward(ctx, grad_y, grad_token):
        xc, weight, bias = ctx.saved_tensors
        bn, lay, run = ctx.bn, ctx.lay, ctx.run
        need_p = (weight is not None and ctx.needs_input_grad[1]) or (bias is not None and ctx.needs_input_grad[2])
        none6 = (None,) * 6
        if grad_y is None and grad_token is None:
            return (None, None, None) + none6
        sl = slice(lay.offset, lay.offset + lay.channels)
        count = run.counts[ctx.idx] * run.world
        nhwc = xc.is_contiguous(memory_format=torch.channels_last) and not xc.is_contiguous() and xc.shape[1] % 4 == 0
        if grad_y is not None and grad_token is not None and not need_p and nhwc:
            g = ops.bn_eval_tap_backward(xc, grad_y, weight, bias, bn.running_mean, bn.running_var, bn.eps, run.mean[sl],
                                         run.gmean[sl], run.gvar[sl], count, relu=ctx.relu,
                                         gscale=grad_token.reshape(1).contiguous())
            return (g, None, None) + none6
        gx = dw = db = None
        if grad_y is not None:
            gx, dw, db = ops.bn_eval_backward(xc, grad_y, weight, bias, bn.running_mean, bn.running_var, bn.eps,
                                              relu=ctx.relu, want_param_grads=need_p)
        if grad_token is not None:
            gx = ops.bn_stats_backward(xc, gx, run.mean[sl], run.gmean[sl], run.gvar[sl], count,
                                       gscale=grad_token.reshape(1).contiguous())
        return (gx, dw if (weight is not None and ctx.needs_input_grad[1]) else None,
                db if (bias is not None and ctx.needs_input_grad[2]) else None) + none6


class PendingTap:
    """What the statistics hook leaves on a fused BatchNorm (``module._pending_tap``) instead of tapping the input
    itself: the module's forward, called next, decides whether tap and BatchNorm run as one node (``fused``) or the
    tap runs on its own in front of whatever the module does (``plain``)."""

    __slots__ = ("mgr", "idx", "run")

    def __init__(self, mgr, idx, run):
        self.mgr, self.idx, self.run = mgr, idx, run

    def plain(self, x):
        x_out, self.run.tokens[self.idx] = _Tap.apply(x, self.mgr, self.idx, self.run)
        return x_out

    def fused(self, x, weight, bias, bn, relu, qact):
        y, self.run.tokens[self.idx] = _TapFusedBN.apply(x, weight, bias, bn, relu, qact, self.mgr, self.idx, self.run)
        return y


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mgr, run, *tokens):
        mgr._reduce_and_evaluate(run)
        return run.loss3[0].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        n = len(ctx.needs_input_grad) - 2
        return (None, None) + (grad_loss,) * n


class _Layer:
    __slots__ = ("module", "offset", "channels")

    def __init__(self, module, offset):
        self.module, self.offset, self.channels = module, offset, module.num_features


class BNStatLoss:
    """BN-statistics matching loss for a whole model.

        bns = BNStatLoss(teacher)            # hooks every BatchNorm module
        out = teacher(images)                # statistics collected during the forward
        loss = ce(out, labels) + 0.1 * bns.loss()
        loss.backward()                      # loss gradient fused into each BN-input gradient

    ``sync=True`` (with an initialised process group): all-reduce the packed sums (one
    NCCL call of ``2 * sum_l C_l`` floats) so that loss and gradients are those of the
    global batch.  After ``loss()``: ``means()``, ``variances()`` give the per-layer
    statistics the reference keeps in ``mean_list`` / ``var_list``; ``parts()`` =
    (mean term, var term).
    """

    def __init__(self, model: nn.Module, bn_types=(nn.modules.batchnorm._BatchNorm,), sync=False,
                 process_group=None):
        self._layers = []
        off = 0
        for m in model.modules():
            if isinstance(m, bn_types):
                if m.running_mean is None:
                    raise RuntimeError("BNStatLoss needs BatchNorm modules that track running statistics")
                self._layers.append(_Layer(m, off))
                off += m.num_features
        if not self._layers:
            raise RuntimeError("BNStatLoss: model has no BatchNorm module")
        self._ctot = off
        self._sync, self._group = sync, process_group
        self._run = None          # the pass being collected
        self._last = None         # the pass loss() was last called on
        self._handles = [lay.module.register_forward_pre_hook(self._make_hook(i))
                         for i, lay in enumerate(self._layers)]

    # ------------------------------------------------------------------ hooks
    def _make_hook(self, idx):
        def hook(module, inputs):
            x = inputs[0]
            if self._run is None:
                self._run = _Pass(len(self._layers), self._ctot, x.device)
            if getattr(module, "_oodfq_accepts_tap", False):
                # a fused BatchNorm (fusion.FusedEvalBN): it picks the tap up in its forward, which runs next
                object.__setattr__(module, "_pending_tap", PendingTap(self, idx, self._run))
                return None
            x_out, self._run.tokens[idx] = _Tap.apply(x, self, idx, self._run)
            return (x_out,) + tuple(inputs[1:])
        # a module that bypasses this BatchNorm's forward (the fused residual tail, fusion._FusedUnitMixin) recognises
        # the hook by this mark and runs the tap itself on the tensor the BatchNorm would have seen
        hook._oodfq_bns_tap = True
        return hook

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def clear(self):
        """Drop a partially collected pass (the reference's ``mean_list.clear()``)."""
        self._run = None

    # ------------------------------------------------------------------ evaluation
    def _packed_running_stats(self, device):
        # Re-packed on every evaluation (two small concatenations): a cache keyed on the buffers' version counters
        # would miss writes made through ``.data`` -- which is exactly how the reference's BN-statistic delta
        # correction rewrites running statistics (trainer_direct.py:292-297) -- and the loss would silently keep
        # matching against the old targets.
        rm = torch.cat([lay.module.running_mean.detach().reshape(-1) for lay in self._layers]).float()
        rv = torch.cat([lay.module.running_var.detach().reshape(-1) for lay in self._layers]).float()
        return rm.to(device), rv.to(device)

    def _reduce_and_evaluate(self, run):
        if self._sync:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(run.sums, op=dist.ReduceOp.SUM, group=self._group)
                run.world = dist.get_world_size(self._group)
        rm, rv = self._packed_running_stats(run.sums.device)
        offs = [lay.offset for lay in self._layers] + [self._ctot]
        counts = [c * run.world for c in run.counts]
        run.loss3, run.mean, run.var, run.gmean, run.gvar = ops.bns_loss(run.sums, rm, rm, rv, offs, counts)

    def loss(self, flavour: str = "trainer") -> torch.Tensor:
        """``sum_l [MSE(mean_l, rm_l) + MSE(var_l, rv_l)] / L`` as a differentiable 0-dim tensor.

        ``flavour``: the reference writes this sum two ways -- ``(sum_l mean_l-term + var_l-term) / L``
        (trainer_direct.py:473-486) and ``sum_l mean-term / L + sum_l var-term / L`` (distill_data.py:252-265).
        The kernel accumulates all terms in fp64 and rounds once, so both are the same number here; the argument
        only keeps the call interchangeable with the oracle's ``StatTap.loss(flavour)``."""
        if flavour not in ("trainer", "distill"):
            raise ValueError(f"BNStatLoss.loss: unknown flavour {flavour!r}")
        run = self._run
        if run is None:
            raise RuntimeError("BNStatLoss.loss(): no forward pass has been collected")
        bad = [i for i, f in enumerate(run.fired) if f != 1]
        if bad:
            self._run = None
            raise RuntimeError(f"BNStatLoss.loss(): BN layers {bad[:8]} fired {[run.fired[i] for i in bad[:8]]} "
                               "times in this pass; each hooked module must run exactly once per loss")
        out = _Loss.apply(self, run, *run.tokens)
        self._run, self._last = None, run
        return out

    def parts(self):
        """(mean term, var term) of the last loss, each already divided by L (distill_data.py:262-264)."""
        return self._last.loss3[1], self._last.loss3[2]

    def means(self):
        return [self._last.mean[lay.offset: lay.offset + lay.channels] for lay in self._layers]

    def variances(self):
        return [self._last.var[lay.offset: lay.offset + lay.channels] for lay in self._layers]
