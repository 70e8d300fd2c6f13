"""BN-statistics matching loss, restated on torch CPU with autograd.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Paths relative to
``/root/reference``.
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


def channel_stats(x):
    """Per-channel batch mean and biased variance of a BN input.

    trainer_direct.py:388-393 and data_generate/distill_data.py:69-73.
    """
    mean = x.mean([0, 2, 3])
    var = x.var([0, 2, 3], unbiased=False)
    return mean, var


def bns_loss_trainer(means, vars_, run_means, run_vars):
    """trainer_direct.py:473-486: sum_l [MSE(mean)+MSE(var)] / L  (weighted 0.1 by the caller)."""
    total = torch.zeros(1)
    for m, v, rm, rv in zip(means, vars_, run_means, run_vars):
        total = total + (F.mse_loss(m, rm) + F.mse_loss(v, rv))
    return total / len(means)


def bns_loss_distill(means, vars_, run_means, run_vars):
    """distill_data.py:252-265: mean term / L + var term / L."""
    lm = torch.zeros(1)
    lv = torch.zeros(1)
    for m, v, rm, rv in zip(means, vars_, run_means, run_vars):
        lm = lm + F.mse_loss(m, rm.detach())
        lv = lv + F.mse_loss(v, rv.detach())
    n = len(means)
    return lm / n + lv / n


def bns_input_grad(x, run_mean, run_var, upstream=1.0):
    """Closed form of d[MSE(mean,rm)+MSE(var,rv)]/dx.

    grad[n,c,h,w] = upstream * ( 4/(C*M) * (var_c-rv_c) * (x-mean_c) + 2/(C*M) * (mean_c-rm_c) ),
    M = N*H*W.  Used to cross-check autograd in the tests.
    """
    n, c, h, w = x.shape
    m = n * h * w
    mean, var = channel_stats(x)
    a = (4.0 / (c * m)) * (var - run_var)
    b = (2.0 / (c * m)) * (mean - run_mean)
    return upstream * (a.view(1, c, 1, 1) * (x - mean.view(1, c, 1, 1)) + b.view(1, c, 1, 1))


class StatTap:
    """Forward hooks on every BatchNorm2d, collecting what the reference hook collects.

    trainer_direct.py:388-397 (registered at :418-423) / distill_data.py:69-78 (:156-158).
    """

    def __init__(self, model, bn_types=(nn.BatchNorm2d,)):
        self.means, self.vars, self.run_means, self.run_vars = [], [], [], []
        self.handles = [m.register_forward_hook(self._hook)
                        for m in model.modules() if isinstance(m, bn_types)]

    def _hook(self, module, inputs, output):
        mean, var = channel_stats(inputs[0])
        self.means.append(mean)
        self.vars.append(var)
        self.run_means.append(module.running_mean)
        self.run_vars.append(module.running_var)

    def clear(self):
        for lst in (self.means, self.vars, self.run_means, self.run_vars):
            lst.clear()

    def loss(self, flavour="trainer"):
        fn = bns_loss_trainer if flavour == "trainer" else bns_loss_distill
        return fn(self.means, self.vars, self.run_means, self.run_vars)

    def remove(self):
        for h in self.handles:
            h.remove()
