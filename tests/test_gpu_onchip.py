"""Single-pass calibrating QuantAct (cooperative kernel: x staged into shared memory by TMA bulk copies and held there
across the grid-wide range reduction) against the two-kernel path: range state and output must be bit-identical.
Round 1 kept the TMA-fed kernel opt-in behind OODFQ_EXPERIMENTAL; it ran on a B200 at the start of round 2
(profiles/r2_pytest_experimental.log, 54 passed) and is the only on-chip kernel since."""
import numpy as np
import pytest
import torch

from oracle import fq_torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def run(x, k, onchip, steps=1):
    from ood_dfq_b200 import ops
    st = [t.to(DEV) for t in (torch.zeros(1), torch.zeros(1), torch.tensor([0.9]), torch.ones(1))]
    y = None
    for _ in range(steps):
        y = ops.act_calib_forward(x, k, *st, onchip=onchip)
    return y, torch.cat(st).cpu().numpy().view(np.int32)


@pytest.mark.parametrize("shape", [(4,), (1, 4, 2, 2), (64, 64, 28, 28), (256, 512, 7, 7), (256, 128, 28, 20), (7, 12, 36, 4),
                                   (148 * 4 + 4,), (148 * 1024 * 4 + 4,), (148 * 1024 * 14 * 4 + 32,), (20 * 1000 * 1000,)])
@pytest.mark.parametrize("k", [2, 4, 8])
def test_onchip_equals_two_kernel_path(shape, k):
    g = torch.Generator().manual_seed(sum(shape) + k)
    x = torch.relu(torch.randn(shape, generator=g) * 1.7).to(DEV)
    a, sa = run(x, k, True, steps=2)
    b, sb = run(x, k, False, steps=2)
    assert np.array_equal(sa, sb)
    assert np.array_equal(a.cpu().numpy().view(np.int32), b.cpu().numpy().view(np.int32))


def test_onchip_matches_the_oracle_module_and_keeps_nan():
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    g = torch.Generator().manual_seed(3)
    ours, ref = qm.QuantAct(4).to(DEV), fq_torch.OracleQuantAct(4)
    for _ in range(3):
        x = torch.relu(torch.randn(8, 16, 12, 12, generator=g) * 1.2)
        y, y_ref = ours(x.to(DEV)), ref(x)
        assert np.array_equal(y.cpu().numpy().view(np.int32), y_ref.numpy().view(np.int32))
        for name in ("x_min", "x_max", "beta_t"):
            assert np.array_equal(getattr(ours, name).cpu().numpy().view(np.int32), getattr(ref, name).numpy().view(np.int32))
    x = torch.randn(2, 4, 8, 8, generator=g)
    x[0, 0, 0, 0] = -0.0
    a, sa = run(x.to(DEV), 4, True)
    b, sb = run(x.to(DEV), 4, False)
    assert np.array_equal(sa, sb) and torch.equal(a.view(torch.int32), b.view(torch.int32))
    x[1, 2, 3, 4] = float("nan")
    a, sa = run(x.to(DEV), 4, True)
    b, sb = run(x.to(DEV), 4, False)
    assert np.array_equal(sa, sb) and torch.isnan(a).all() and torch.isnan(b).all()     # NaN range -> NaN everywhere


def test_odd_sizes_and_wide_bit_widths_take_the_other_path():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 5, 7, generator=g).to(DEV)            # numel % 4 != 0
    a, sa = run(x, 4, True)
    b, sb = run(x, 4, False)
    assert np.array_equal(sa, sb) and torch.equal(a, b)
    x = torch.randn(4, 8, generator=g).to(DEV)               # k > 8: no table
    a, sa = run(x, 12, True)
    b, sb = run(x, 12, False)
    assert np.array_equal(sa, sb) and torch.equal(a, b)
