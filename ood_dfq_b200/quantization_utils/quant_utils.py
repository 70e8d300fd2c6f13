"""Drop-in for the reference's ``quantization_utils/quant_utils.py`` on B200.

Same public names, argument meaning and return conventions as the reference module
(file:line cited per symbol, relative to the reference tree); every tensor operation
is one launch of a hand-written sm_100a kernel through the C ABI in
``include/oodfq_b200.h``.  CUDA fp32 tensors only -- CPU tensors raise, there is no
fallback path.

The star-export surface matters: ``quant_modules`` re-exports this module with
``import *`` and the reference's callers rely on the globals that leak through it
(``np`` in main_direct.py:183-195), so ``math``, ``np``, ``torch``, ``Function`` and
``Variable`` are module globals here as well.
"""
import math  # noqa: F401  (part of the star-export surface)

import numpy as np  # noqa: F401
import torch
from torch.autograd import Function, Variable  # noqa: F401

from .. import _native as _N
from .. import ops as _ops


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """L_p distance used by the MSE range search (reference quant_utils.py:26-33)."""
    err = (pred - tgt).abs().pow(p)
    return err.sum(1).mean() if reduction == 'none' else err.mean()


def find_MSESmallest(x, k, x_min=None, x_max=None):
    """Fake-quantise ``x`` with a candidate range (reference quant_utils.py:36-47): one fused launch."""
    return _ops.fake_quant(x, k, x_min, x_max)


def clamp(input, min, max, inplace=False):
    """Saturate to [min, max] (reference quant_utils.py:49-58).

    Not on the hot path any more: the fused kernels clamp the codes in registers.
    """
    if inplace:
        return input.clamp_(min, max)
    return torch.clamp(input, min, max)


def _as_param(v, like):
    if isinstance(v, torch.Tensor):
        return v
    return torch.tensor([float(v)], dtype=torch.float32, device=like.device)


def _elementwise(input, scale, zero_point, inplace, mode, symmetric):
    scale = _as_param(scale, input)
    zero_point = _as_param(zero_point, input)
    if inplace:
        if not (input.is_contiguous() or (input.dim() == 4 and input.is_contiguous(memory_format=torch.channels_last))):
            raise RuntimeError("ood_dfq_b200: inplace=True needs a dense tensor")
        _ops.elementwise(input, scale, zero_point, 8, mode, symmetric=symmetric, params_given=True, out=input)
        return input
    return _ops.elementwise(input, scale, zero_point, 8, mode, symmetric=symmetric, params_given=True)


def linear_quantize(input, scale, zero_point, inplace=False):
    """``round(scale * input - zero_point)``, scale/zero-point per dim-0 row (reference quant_utils.py:61-81)."""
    return _elementwise(input, scale, zero_point, inplace, _N.MODE_QUANTIZE, False)


def linear_dequantize(input, scale, zero_point, inplace=False):
    """``(input + zero_point) / scale`` with a true division (reference quant_utils.py:84-104)."""
    return _elementwise(input, scale, zero_point, inplace, _N.MODE_DEQUANTIZE, False)


def linear_quantize_DSG(input, scale, zero_point, inplace=False):
    """``round(scale * input)``; zero-point ignored (reference quant_utils.py:192-212)."""
    return _elementwise(input, scale, zero_point, inplace, _N.MODE_QUANTIZE, True)


def linear_dequantize_DSG(input, scale, zero_point, inplace=False):
    """``input / scale``; zero-point ignored (reference quant_utils.py:215-235)."""
    return _elementwise(input, scale, zero_point, inplace, _N.MODE_DEQUANTIZE, True)


def _range_params(num_bits, saturation_min, saturation_max, integral_zero_point, signed):
    if not (integral_zero_point and signed):
        # the reference's callers never pass anything else; keep the contract explicit
        raise NotImplementedError("ood_dfq_b200: only integral_zero_point=True, signed=True is implemented")
    return _ops.quant_params(num_bits, saturation_min, saturation_max)


def asymmetric_linear_quantization_params(num_bits, saturation_min, saturation_max,
                                          integral_zero_point=True, signed=True):
    """(scale, zero_point) of a range (reference quant_utils.py:107-128).

    scale = reciprocal(clamp(max - min, 1e-8)) * (2^k - 1); zero_point = round(scale * min) + 2^(k-1).
    """
    return _range_params(num_bits, saturation_min, saturation_max, integral_zero_point, signed)


def symmetric_linear_quantization_params_DSG(num_bits, saturation_min, saturation_max,
                                             integral_zero_point=True, signed=True):
    """Same formula as the asymmetric helper (reference quant_utils.py:238-259)."""
    return _range_params(num_bits, saturation_min, saturation_max, integral_zero_point, signed)


def _ste_forward(x, k, x_min, x_max, symmetric):
    if x_min is None or x_max is None:
        raise RuntimeError("ood_dfq_b200: x_min and x_max are required (the reference has no default either)")
    return _ops.fake_quant(x, k, x_min, x_max, symmetric=symmetric)


class AsymmetricQuantFunction(Function):
    """Fake-quantise with a given range; straight-through backward (reference quant_utils.py:131-161).

    forward(x, k, x_min, x_max): ``x_min`` / ``x_max`` hold one range (activations) or one
    range per dim-0 row (weights).  backward hands ``grad_output`` through untouched and
    gives no gradient to the range -- exactly the reference's identity STE, no clip mask.
    """

    @staticmethod
    def forward(ctx, x, k, x_min=None, x_max=None):
        return _ste_forward(x, k, x_min, x_max, False)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None, None


class SymmetricQuantFunction_DSG(Function):
    """``round(scale*x)`` / ``q/scale`` variant, same STE (reference quant_utils.py:262-290)."""

    @staticmethod
    def forward(ctx, x, k, x_min=None, x_max=None):
        return _ste_forward(x, k, x_min, x_max, True)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None, None
