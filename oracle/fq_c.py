"""numpy front-end of the plain-C restatement (oracle/fq_oracle.c).  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from . import build_c

_lib = None
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.fqc_params.argtypes = [C.c_int, C.c_float, C.c_float, _fp, _fp]
        _lib.fqc_fake_quant.argtypes = [_fp, _fp, _fp, C.c_size_t, C.c_size_t, C.c_int, _fp, _fp, C.c_int]
        _lib.fqc_minmax.argtypes = [_fp, C.c_size_t, _fp, _fp]
        _lib.fqc_row_ranges.argtypes = [_fp, C.c_size_t, C.c_size_t, C.c_int, _fp, _fp]
        _lib.fqc_range_update.argtypes = [_fp, C.c_float, C.c_float, C.c_float, C.c_int]
        _lib.fqc_channel_stats.argtypes = [_fp, C.c_size_t, C.c_size_t, C.c_size_t, _dp, _dp]
        _lib.fqc_mse_search.argtypes = [_fp, C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_float, _fp, _fp]
        _lib.fqc_mse_search.restype = C.c_int
        for f in ("fqc_params", "fqc_fake_quant", "fqc_minmax", "fqc_row_ranges", "fqc_range_update", "fqc_channel_stats"):
            getattr(_lib, f).restype = None
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_fp)


def params(k, lo, hi):
    s, z = C.c_float(), C.c_float()
    lib().fqc_params(int(k), float(lo), float(hi), C.byref(s), C.byref(z))
    return np.float32(s.value), np.float32(z.value)


def fake_quant(x, k, lo, hi, symmetric=False):
    """Returns (y, codes); lo/hi scalars or per-row arrays."""
    x = _f(x)
    lo, hi = _f(np.reshape(lo, -1)), _f(np.reshape(hi, -1))
    y, codes = np.empty_like(x), np.empty_like(x)
    lib().fqc_fake_quant(_p(x), _p(y), _p(codes), x.size, lo.size, int(k), _p(lo), _p(hi), int(symmetric))
    return y, codes


def minmax(x):
    x = _f(x)
    a, b = C.c_float(), C.c_float()
    lib().fqc_minmax(_p(x), x.size, C.byref(a), C.byref(b))
    return np.float32(a.value), np.float32(b.value)


def row_ranges(w, symmetric=False):
    w = _f(w)
    rows = w.shape[0]
    lo, hi = np.empty(rows, np.float32), np.empty(rows, np.float32)
    lib().fqc_row_ranges(_p(w), rows, w.size // rows, int(symmetric), _p(lo), _p(hi))
    return lo, hi


def range_update(state, beta, data_min, data_max, symmetric=False):
    st = _f(state).copy()
    lib().fqc_range_update(_p(st), float(beta), float(data_min), float(data_max), int(symmetric))
    return st


def channel_stats(x):
    x = _f(x)
    n, c, h, w = x.shape
    mean, var = np.empty(c, np.float64), np.empty(c, np.float64)
    lib().fqc_channel_stats(_p(x), n, c, h * w, mean.ctypes.data_as(_dp), var.ctypes.data_as(_dp))
    return mean, var


def mse_search(x, k, state, beta=0.9, steps=80, p=2.4):
    """QuantAct_MSE's clip search + plain EMA.  Returns (new state [x_min, x_max, beta_t], scores, kept index)."""
    x = _f(x).reshape(-1)
    st = _f(state).copy()
    scores = np.empty(steps, np.float32)
    keep = lib().fqc_mse_search(_p(x), x.size, int(k), int(steps), float(p), float(np.float32(beta)), _p(st), _p(scores))
    return st, scores, int(keep)
