"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool).

Every output tensor the launch wrappers allocate is carved out of a larger buffer pre-filled with a sentinel;
after the kernels ran, the guard bands on both sides must be untouched.  Shapes are chosen to hit the ragged
tails: numel % 8 != 0, 7x7 / 3x5 planes, C not a multiple of the group size, rows longer than a CTA, ...
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 512            # elements on each side (keeps 32-byte alignment for every dtype used)
SENT = {torch.float32: 1234567.0, torch.float64: 1234567.0, torch.int8: 77, torch.uint8: 77}


class Guard:
    def __init__(self):
        self.bases = []

    def alloc(self, shape, strides, dtype, device):
        n = 0 if any(s == 0 for s in shape) else 1 + sum((s - 1) * st for s, st in zip(shape, strides))
        base = torch.full((n + 2 * PAD,), SENT[dtype], dtype=dtype, device=device)
        self.bases.append((base, n))
        return torch.as_strided(base, tuple(shape), tuple(strides), PAD)

    def empty_like(self, t, dtype=None, **kw):
        dtype = dtype or t.dtype
        if dtype not in SENT or not t.is_cuda:
            return _REAL_EMPTY_LIKE(t, dtype=dtype, **kw)
        dense = t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))
        strides = t.stride() if dense else torch.empty(t.shape, device="meta").stride()
        return self.alloc(t.shape, strides, dtype, t.device)

    def empty(self, *size, dtype=torch.float32, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        if dtype not in SENT or device is None or torch.device(device).type != "cuda":
            return _REAL_EMPTY(*size, dtype=dtype, device=device, **kw)
        return self.alloc(shape, torch.empty(shape, device="meta").stride(), dtype, device)

    def check(self):
        assert self.bases, "guard allocated nothing"
        for base, n in self.bases:
            s = SENT[base.dtype]
            assert bool((base[:PAD] == s).all()) and bool((base[PAD + n:] == s).all()), "write outside an output tensor"


_REAL_EMPTY_LIKE, _REAL_EMPTY = torch.empty_like, torch.empty


@pytest.fixture
def guard(monkeypatch):
    from ood_dfq_b200 import ops
    g = Guard()

    class _T:                                   # a torch facade for the ops module only
        def __getattr__(self, name):
            return getattr(torch, name)
    facade = _T()
    facade.empty_like = g.empty_like
    facade.empty = g.empty
    monkeypatch.setattr(ops, "torch", facade)
    yield g
    g.check()


def rnd(shape, seed=0, relu=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    return (torch.relu(x) if relu else x).to(DEV)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1023, 8191, 8193, 70001])
def test_flat_kernels_stay_in_bounds(guard, n):
    from ood_dfq_b200 import ops
    x = rnd((n,), n)
    lo, hi = torch.zeros(1, device=DEV), torch.ones(1, device=DEV)
    ops.fake_quant(x, 4, lo, hi)
    ops.fake_quant(x, 4, lo, hi, codes=True)
    ops.fake_quant(x, 4, -hi, hi, symmetric=True)
    st = [torch.zeros(1, device=DEV), torch.zeros(1, device=DEV), torch.full((1,), 0.9, device=DEV), torch.ones(1, device=DEV)]
    ops.act_calib_forward(x, 4, *st)
    ops.minmax(x)


@pytest.mark.parametrize("shape", [(3, 5), (9, 27), (10, 64), (17, 147), (5, 1025), (3, 4608), (2, 5001), (33, 7, 3, 3)])
def test_weight_kernel_stays_in_bounds(guard, shape):
    from ood_dfq_b200 import ops
    w = rnd(shape, sum(shape), relu=False) * 0.05
    ops.weight_fq_multi([w, w * 2], [4, 2], [False, True], want_range=True, want_codes=True)
    ops.weight_fq_multi([w], [4], [False])
    lo = torch.full((shape[0],), -0.1, device=DEV)
    ops.fake_quant(w, 4, lo, -lo)               # the per-row generic path


@pytest.mark.parametrize("shape", [(3, 5, 7, 7), (2, 3, 3, 5), (5, 130, 7, 7), (4, 16, 33, 35), (2, 6, 70, 70),
                                   (3, 64, 56, 56), (9, 512, 4, 4), (1, 1, 1, 1), (2, 8, 1, 3)])
def test_per_channel_kernels_stay_in_bounds(guard, shape):
    from ood_dfq_b200 import ops
    c = shape[1]
    x, gy = rnd(shape, 1, relu=False), rnd(shape, 2, relu=False)
    w, b = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV)
    rm, rv = torch.randn(c, device=DEV) * 0.1, torch.rand(c, device=DEV) + 0.5
    lo, hi = torch.zeros(1, device=DEV), torch.ones(1, device=DEV) * 2
    cnt = float(x.numel() // c)
    sums = ops.bn_stats_forward(x, rm)
    ops.bn_stats_forward(x, rm, fq=(4, lo, hi))
    mean, var = ops.bn_stats_finalize(sums, rm, cnt)
    ops.bn_stats_backward(x, gy, mean, w, b, cnt)
    ops.bn_stats_backward(x, None, mean, w, b, cnt)
    for xx, gg in ((x, gy), (x.contiguous(memory_format=torch.channels_last), gy.contiguous(memory_format=torch.channels_last))):
        ops.bn_eval_forward(xx, w, b, rm, rv, 1e-5, relu=True, fq=(4, lo, hi), want_z=True)
        ops.bn_eval_forward(xx, w, b, rm, rv, 1e-5)
        ops.bn_eval_backward(xx, gg, w, b, rm, rv, 1e-5, relu=True)
        ops.bn_eval_backward(xx, gg, w, b, rm, rv, 1e-5, relu=False, want_param_grads=False)


def _bn(c, seed):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(c, generator=g) + 0.5).to(DEV), torch.randn(c, generator=g).to(DEV) * 0.3,
            (torch.randn(c, generator=g) * 0.1).to(DEV), (torch.rand(c, generator=g) + 0.5).to(DEV))


@pytest.mark.parametrize("shape", [(3, 64, 30, 30), (2, 16, 9, 11), (5, 8, 7, 7), (2, 128, 5, 6), (1, 4, 1, 1), (7, 12, 13, 3),
                                   (2, 64, 112, 112), (150, 4, 6, 6)])
def test_fused_stem_kernels_stay_in_bounds(guard, shape):
    """Both stem forwards (TMA-staged ring and register kernel) and both backwards write out / idx / xhat / grad_x only
    inside their tensors: odd extents, one-window rows, more items than CTAs, rows that wrap the ring."""
    from ood_dfq_b200 import ops
    x = rnd(shape, 3, relu=False).contiguous(memory_format=torch.channels_last)
    w, b, rm, rv = _bn(shape[1], 4)
    lo, hi = torch.zeros(1, device=DEV), torch.ones(1, device=DEV) * 2
    for fq in (None, (4, lo, hi)):
        for reg in (False, True):
            out, idx, xhat = ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, register_kernel=reg)
            ops.bn_pool_forward(x, w, b, rm, rv, 1e-5, fq=fq, want_xhat=False, register_kernel=reg)
        go = torch.randn_like(out)
        ops.bn_pool_backward(go, idx, xhat, x.shape, w, b, rm, rv, 1e-5)
        ops.bn_pool_backward(go, idx, None, x.shape, w, b, rm, rv, 1e-5, want_param_grads=False, grad_out2=go)


@pytest.mark.parametrize("shape", [(3, 64, 7, 7), (2, 16, 9, 11), (5, 8, 1, 1), (2, 512, 7, 7), (33, 4, 5, 3), (2, 1024, 2, 2),
                                   (9, 16, 32, 32)])
def test_residual_tail_and_pool_kernels_stay_in_bounds(guard, shape):
    from ood_dfq_b200 import ops
    c = shape[1]
    x1 = rnd(shape, 5, relu=False).contiguous(memory_format=torch.channels_last)
    r = rnd(shape, 6).contiguous(memory_format=torch.channels_last)
    gy = rnd(shape, 7, relu=False).contiguous(memory_format=torch.channels_last)
    ge = torch.randn(shape[:2], device=DEV)
    bn1, bn2 = _bn(c, 8) + (1e-5,), _bn(c, 9) + (1e-5,)
    lo, hi = torch.zeros(1, device=DEV), torch.ones(1, device=DEV) * 2
    for b2 in (None, bn2):
        y, e, mask = ops.res_tail_forward(x1, r, bn1, b2, fq=(4, lo, hi), want_energy=True, want_mask=True)
        ops.res_tail_forward(x1, r, bn1, b2)
        ops.res_tail_backward(gy, ge, x1, r, bn1, b2)
        ops.res_tail_backward(gy, ge, x1, r if b2 is not None else None, bn1, b2, mask=mask, grad_y2=gy)
        ops.res_tail_backward(gy, None, None, None, bn1, b2, want_param_grads=False, mask=mask)
    if shape[2] * shape[3] > 1:
        p = ops.global_avgpool_forward(r)
        ops.global_avgpool_backward(torch.randn_like(p), r.shape)
    _, m = ops.bn_eval_forward(x1, *bn1[:4], 1e-5, relu=True, fq=(4, lo, hi), want_mask=True)
    if m is not None:                       # (a 1x1 plane is NCHW-contiguous too and takes the kernels without a mask)
        ops.bn_eval_backward(None, gy, *bn1[:4], 1e-5, relu=True, want_param_grads=False, mask=m)


@pytest.mark.parametrize("shape", [(3, 3, 32, 32), (2, 3, 224, 224), (5, 1, 28, 28), (2, 4, 10, 6)])
def test_stem_relayout_stays_in_bounds(guard, shape):
    from ood_dfq_b200 import ops
    x = rnd(shape, 10, relu=False).contiguous(memory_format=torch.channels_last)
    if not ops.s2d_stem_supported(x, 3):
        pytest.skip("shape not taken by the re-layout kernel")
    for cpad in (None, 16):
        xs = ops.s2d_stem_forward(x, 3, cpad=cpad)
        ops.s2d_stem_backward(torch.randn_like(xs), x.shape, 3)
