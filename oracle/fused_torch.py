"""Eval-mode BatchNorm -> ReLU -> QuantAct, restated on torch CPU (oracle for the fused kernels).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference never fuses these: it runs
``nn.BatchNorm2d`` in eval() (trainer_direct.py:411-412) followed by the ``nn.Sequential(ReLU, QuantAct)``
that ``quantize_model`` builds (main_direct.py:464-465).  The affine is evaluated in fp64 here so that the
kernel's fp32 result can be checked to half an ulp; the fake-quant of that fp32 value must then be bit-exact.
"""
import torch
import torch.nn.functional as F

from . import fq_torch


def bn_eval_affine64(x, weight, bias, running_mean, running_var, eps):
    """z = (x - rm) / sqrt(rv + eps) * w + b in float64 (the mathematical value of eval-mode BatchNorm)."""
    shape = (1, -1, 1, 1)
    w = torch.ones_like(running_mean) if weight is None else weight
    b = torch.zeros_like(running_mean) if bias is None else bias
    inv = (running_var.double() + eps).rsqrt()
    return (x.double() - running_mean.double().view(shape)) * (inv * w.double()).view(shape) + b.double().view(shape)


def bn_relu_quant(x, weight, bias, running_mean, running_var, eps, k, lo, hi, relu=True):
    """The unfused module chain, exactly as torch runs it on CPU (differentiable)."""
    z = F.batch_norm(x, running_mean, running_var, weight, bias, False, 0.0, eps)
    if relu:
        z = F.relu(z)
    return fq_torch.fake_quant_ste(z, k, lo, hi) if k else z


def residual_tail(x1, r, bn1, bn2=None, k=0, lo=None, hi=None):
    """Tail of a residual unit exactly as the unfused modules run it on CPU (differentiable):
    ``QuantAct(ReLU(BN1(x1) + id))`` with ``id = r`` or ``BN2(r)`` -- pytorchcv ResUnit.forward / reference
    models.py:40-47 after quantize_model (main_direct.py:464-465) -- plus the trainer's channel attention input
    ``BN1(x1).pow(2).mean([2, 3])`` (trainer_direct.py:382-383).  ``bn = (weight, bias, running_mean,
    running_var, eps)``.  Returns ``(y, energy, pre_activation)``."""
    w1, b1, rm1, rv1, eps1 = bn1
    z1 = F.batch_norm(x1, rm1, rv1, w1, b1, False, 0.0, eps1)
    if bn2 is not None:
        w2, b2, rm2, rv2, eps2 = bn2
        ident = F.batch_norm(r, rm2, rv2, w2, b2, False, 0.0, eps2)
    else:
        ident = r
    s = z1 + ident
    y = F.relu(s)
    if k:
        y = fq_torch.fake_quant_ste(y, k, lo, hi)
    return y, z1.pow(2).mean([2, 3]), s
