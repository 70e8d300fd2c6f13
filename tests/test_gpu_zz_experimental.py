"""Kernels that were written without GPU time and are therefore OPT-IN: they are not on any default path and
these tests only run with OODFQ_EXPERIMENTAL=1 (first thing to do with GPU minutes, see tools/gpu_session.sh).

* act_calib_onchip_tma_kernel (csrc/fq_calib.cu): the single-pass calibrating QuantAct with its shared-memory tile
  filled by TMA bulk copies -- must be bit-identical to the register-staged kernel and to the two-kernel path.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("OODFQ_EXPERIMENTAL"), reason="opt-in: set OODFQ_EXPERIMENTAL=1")]

DEV = "cuda:0"


def run(x, k, onchip, steps=1):
    from ood_dfq_b200 import ops
    st = [t.to(DEV) for t in (torch.zeros(1), torch.zeros(1), torch.tensor([0.9]), torch.ones(1))]
    y = None
    for _ in range(steps):
        y = ops.act_calib_forward(x, k, *st, onchip=onchip)
    return y, torch.cat(st).cpu().numpy().view(np.int32)


# (148*1024*14 + 8)*4 elements: one vector more than the chip holds; 20M elements: the through-L2 remainder
@pytest.mark.parametrize("shape", [(4,), (1, 4, 2, 2), (64, 64, 28, 28), (256, 512, 7, 7), (256, 128, 28, 20), (7, 12, 36, 4),
                                   (148 * 4 + 4,), (148 * 1024 * 4 + 4,), (148 * 1024 * 14 * 4 + 32,), (20 * 1000 * 1000,)])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_tma_variant_is_bit_identical(shape, k):
    g = torch.Generator().manual_seed(sum(shape) % 1000 + k)
    x = torch.relu(torch.randn(shape, generator=g) * 1.7).to(DEV)
    a, sa = run(x, k, "tma", steps=2)
    b, sb = run(x, k, True, steps=2)
    c, sc = run(x, k, False, steps=2)
    assert np.array_equal(sa, sb) and np.array_equal(sa, sc)
    assert torch.equal(a.view(torch.int32), b.view(torch.int32)) and torch.equal(a.view(torch.int32), c.view(torch.int32))


def test_tma_variant_keeps_nan_and_negative_zero():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 8, 8, generator=g)
    x[0, 0, 0, 0] = -0.0
    a, sa = run(x.to(DEV), 4, "tma")
    b, sb = run(x.to(DEV), 4, False)
    assert np.array_equal(sa, sb) and torch.equal(a.view(torch.int32), b.view(torch.int32))
    x[1, 2, 3, 4] = float("nan")
    a, sa = run(x.to(DEV), 4, "tma")
    b, sb = run(x.to(DEV), 4, False)
    assert np.array_equal(sa, sb) and torch.isnan(a).all() and torch.isnan(b).all()
