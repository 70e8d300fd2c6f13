"""Model surgery: the consumer side of the quantisation API, restated for benchmarks and tests.

The reference does this inside ``ExperimentDesign`` (main_direct.py:444-516), which cannot
be imported (pyhocon / pytorchcv are absent).  Same rules, exact-type dispatch included:

* ``nn.Conv2d``  -> ``Quant_Conv2d(weight_bit).set_param(conv)``
* ``nn.Linear``  -> ``Quant_Linear(weight_bit).set_param(linear)``
* ``nn.ReLU`` / ``nn.ReLU6`` -> ``nn.Sequential(relu, QuantAct(activation_bit))``
* ``nn.Sequential`` -> rebuilt from converted children (child names become indices)
* anything else -> deep copy whose module attributes (names without 'norm') are converted.

``namespace`` supplies the classes; it defaults to this package's CUDA modules, tests pass
the CPU oracle's classes to build the twin model.
"""
from __future__ import annotations

import copy

from torch import nn


def _default_namespace():
    from .quantization_utils import quant_modules
    return quant_modules


def _cls(namespace, *names):
    for n in names:
        if hasattr(namespace, n):
            return getattr(namespace, n)
    raise AttributeError(f"{namespace} has none of {names}")


def quantize_model(model: nn.Module, weight_bit: int, act_bit: int, namespace=None) -> nn.Module:
    ns = namespace or _default_namespace()
    conv_cls = _cls(ns, "Quant_Conv2d", "OracleQuantConv2d")
    lin_cls = _cls(ns, "Quant_Linear", "OracleQuantLinear")
    act_cls = _cls(ns, "QuantAct", "OracleQuantAct")

    def convert(m):
        t = type(m)
        if t is nn.Conv2d:
            q = conv_cls(weight_bit=weight_bit)
            q.set_param(m)
            return q
        if t is nn.Linear:
            q = lin_cls(weight_bit=weight_bit)
            q.set_param(m)
            return q
        if t is nn.ReLU or t is nn.ReLU6:
            return nn.Sequential(m, act_cls(activation_bit=act_bit))
        if t is nn.Sequential:
            return nn.Sequential(*[convert(c) for _, c in m.named_children()])
        clone = copy.deepcopy(m)
        for name, child in m.named_children():
            if 'norm' not in name:
                setattr(clone, name, convert(child))
        return clone

    return convert(model)


def _act_modules(model, namespace):
    ns = namespace or _default_namespace()
    act_cls = _cls(ns, "QuantAct", "OracleQuantAct")
    return [m for m in model.modules() if type(m) is act_cls]


def freeze_model(model: nn.Module, namespace=None):
    """Stop range tracking in every QuantAct (main_direct.py:486-500)."""
    for m in _act_modules(model, namespace):
        m.fix()
    return model


def unfreeze_model(model: nn.Module, namespace=None):
    """Resume range tracking (main_direct.py:502-516)."""
    for m in _act_modules(model, namespace):
        m.unfix()
    return model
