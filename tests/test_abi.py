"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and
exports every symbol include/oodfq_b200.h declares; the Python mirror exposes the reference's names."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oodfq_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oodfq_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from ood_dfq_b200 import build
    return build.build()          # nvcc cross-compiles here; no-op when up to date


def test_header_declares_what_python_binds():
    from ood_dfq_b200 import _native
    assert declared_functions() == sorted(_native.SIGNATURES)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    from ood_dfq_b200 import _native
    loaded = _native.load()
    assert loaded.oodfq_abi_version() == _native.ABI_VERSION
    assert loaded.oodfq_workspace_bytes() > 0
    assert ctypes.sizeof(_native.WeightDesc) == 5 * 8 + 2 * 8 + 2 * 4


def test_bad_arguments_return_error_codes_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from ood_dfq_b200 import _native
    lib = _native.load()
    assert lib.oodfq_fq_forward(None, None, None, 16, None, None, 1, 4, 0, 0, None) == -1
    assert b"null pointer" in lib.oodfq_last_error()
    assert lib.oodfq_fq_forward(1, 1, None, 16, 1, 1, 1, 99, 0, 0, None) == -1
    assert b"k=99" in lib.oodfq_last_error()
    assert lib.oodfq_fq_forward(1, 1, None, 0, 1, 1, 1, 4, 0, 0, None) == 0      # empty tensor: no launch
    assert lib.oodfq_minmax(1, 0, 1, 1, None) == -1
    assert lib.oodfq_bn_stats_forward(1, 0, 4, 4, None, 1, None, None, None, 0, 0, 1, None) == -1
    with pytest.raises(RuntimeError, match="null pointer"):
        _native.check(lib.oodfq_quant_params(None, None, None, None, 1, 4, None), "quant_params")
    # batch assembly: (images, M, C_in, H, W, index, boxes, flips, out, N, C_out, OH, OW, flags, stream)
    crop = lib.oodfq_crop_resize_flip
    assert crop(1, 4, 3, 8, 8, 1, 1, 1, 1, 0, 3, 8, 8, 0, None) == 0               # empty batch: no launch
    assert crop(None, 4, 3, 8, 8, 1, 1, 1, 1, 2, 3, 8, 8, 0, None) == -1 and b"null pointer" in lib.oodfq_last_error()
    assert crop(1, 4, 3, 8, 8, 1, 1, 1, 1, 2, 1, 8, 8, 0, None) == -1 and b"channels 3 -> 1" in lib.oodfq_last_error()
    assert crop(1, 0, 3, 8, 8, 1, 1, 1, 1, 2, 3, 8, 8, 0, None) == -1 and b"empty image set" in lib.oodfq_last_error()
    assert crop(1, 4, 3, 8, 20000, 1, 1, 1, 1, 2, 3, 8, 8, 0, None) == -1 and b"sides above" in lib.oodfq_last_error()
    assert crop(1, 4, 3, 8, 8, 1, 1, 1, 1, 40000, 3, 8192, 8192, 0, None) == -1 and b"2^31" in lib.oodfq_last_error()
    back = lib.oodfq_crop_resize_flip_backward
    assert back(1, None, 4, 3, 8, 8, 1, 1, 1, 2, 3, 8, 8, 0, None) == -1 and b"null pointer" in lib.oodfq_last_error()
    assert back(1, 1, 4, 1, 8, 8, 1, 1, 1, 2, 2, 8, 8, 0, None) == -1 and b"channels 1 -> 2" in lib.oodfq_last_error()
    assert back(1, 1, 4, 3, 8, 8, 1, 1, 1, 0, 3, 8, 8, 0, None) == 0


def test_mirror_exports_reference_names():
    import ood_dfq_b200
    ood_dfq_b200.install()
    ns = {}
    exec("from quantization_utils.quant_modules import *", ns)
    for name in ["QuantAct", "QuantAct_MSE", "Quant_Linear", "Quant_Conv2d", "QuantAct_DSG", "QuantLinear_DSG",
                 "QuantConv2d_DSG", "lp_loss", "find_MSESmallest", "clamp", "linear_quantize", "linear_dequantize",
                 "asymmetric_linear_quantization_params", "AsymmetricQuantFunction", "linear_quantize_DSG",
                 "linear_dequantize_DSG", "symmetric_linear_quantization_params_DSG", "SymmetricQuantFunction_DSG",
                 "np", "torch", "nn", "F", "Module", "Parameter", "math", "time", "sys", "Function", "Variable"]:
        assert name in ns, name
    q = ns["QuantAct"](4)
    assert sorted(q.state_dict()) == ["beta", "beta_t", "x_max", "x_min"]
    assert all(v.shape == (1,) and v.dtype == torch.float32 for v in q.state_dict().values())
    assert repr(q) == ("QuantAct(activation_bit=4, full_precision_flag=False, running_stat=True, "
                       "Act_min: 0.00, Act_max: 0.00)")
    q.fix()
    assert q.running_stat is False
    q.unfix()
    assert q.running_stat is True
    conv = torch.nn.Conv2d(3, 8, 3, stride=2, padding=1)
    qc = ns["Quant_Conv2d"](weight_bit=3)
    qc.set_param(conv)
    assert repr(qc) == "(Quant_Conv2d() weight_bit=3, full_precision_flag=False)"
    assert sorted(n for n, _ in qc.named_parameters()) == ["bias", "weight"]
    assert (qc.in_channels, qc.out_channels, qc.stride, qc.padding, qc.groups) == (3, 8, (2, 2), (1, 1), 1)
    assert qc.weight.data_ptr() != conv.weight.data_ptr() and torch.equal(qc.weight, conv.weight)
    lin = ns["Quant_Linear"](weight_bit=4)
    lin.set_param(torch.nn.Linear(5, 2, bias=False))
    assert lin.bias is None and (lin.in_features, lin.out_features) == (5, 2)


def test_no_cpu_fallback():
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qm.QuantAct(4)(torch.randn(1, 2, 3, 3))
    conv = qm.Quant_Conv2d(4)
    conv.set_param(torch.nn.Conv2d(2, 2, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.randn(1, 2, 3, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ood_dfq_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
