"""ood_dfq_b200 -- B200-native (sm_100a) fake-quantisation hot path of OOD-DFQ.

Layout
    csrc/                  hand-written CUDA kernels + the C ABI (include/oodfq_b200.h)
    _native.py, ops.py     ctypes binding and tensor-level launch wrappers
    quantization_utils/    drop-in mirror of the reference's operator API
    bns.py                 BN-statistics matching loss (fused forward / loss / backward)
    fusion.py              eval-mode BatchNorm (+ ReLU + QuantAct) fusion pass
    dist.py                data-parallel glue: packed range and BN-sum all-reduces
    step.py                host code of the QAT / distillation iteration, flat gradients, CUDA-graph replay
    surgery.py, nets.py    host-side consumers used by the benchmark (model surgery, carrier networks)
    hocon.py               reader for the reference's config/*.hocon files
    shards.py, augment.py  the step's input side: pickle-shard reader, device-side crop / resize / flip batches

``install()`` makes ``from quantization_utils.quant_modules import *`` (main_direct.py:21,
trainer_direct.py:19) resolve to this implementation.
"""
import sys

__version__ = "0.1.0"


def install():
    """Register the mirror as top-level ``quantization_utils`` (call before the reference imports it)."""
    from . import quantization_utils as pkg
    from .quantization_utils import quant_modules, quant_utils
    sys.modules["quantization_utils"] = pkg
    sys.modules["quantization_utils.quant_utils"] = quant_utils
    sys.modules["quantization_utils.quant_modules"] = quant_modules
    return pkg
