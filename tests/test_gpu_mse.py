"""QuantAct_MSE range search as one scoring pass (csrc/fq_mse.cu) against the CPU oracle's 80-iteration loop."""
import numpy as np
import pytest
import torch

from oracle import fq_torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def oracle_scores(x, k, steps=80, p=2.4):
    lo, hi = x.min(), x.max()
    out = []
    for i in range(steps):
        f = 1.0 - (i * 0.01)
        out.append(fq_torch.lp_loss(x, fq_torch.fake_quant(x, k, lo * f, hi * f), p=p, reduction="all").item())
    return np.array(out, dtype=np.float64)


@pytest.mark.parametrize("shape", [(2, 4, 6, 6), (8, 16, 14, 14), (4, 64, 28, 28), (1, 1, 1, 1), (3, 5, 7, 9), (1, 3, 1, 4099)])
@pytest.mark.parametrize("k", [2, 4, 5, 8])
@pytest.mark.parametrize("post_relu", [True, False])
def test_scores_and_choice_match_the_reference_loop(shape, k, post_relu):
    from ood_dfq_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) * 31 + k + post_relu)
    x = torch.randn(shape, generator=g) * 1.5
    if post_relu:
        x = torch.relu(x)
    ref = oracle_scores(x, k)
    state = [t.to(DEV) for t in (torch.zeros(1), torch.zeros(1), torch.tensor([0.9]), torch.ones(1))]
    cur = torch.empty(2, device=DEV)
    scores, chosen = ops.act_mse_search(x.to(DEV), k, *state, cur_min=cur[0:1], cur_max=cur[1:2], debug=True)
    got = scores.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-12)
    assert cur[0].item() == x.min().item() and cur[1].item() == x.max().item()
    # the kept candidate: the reference's first strict minimum, unless two scores tie to within the rounding of
    # the reductions -- then either of the tied candidates is acceptable
    best = 1e10
    keep = 0
    for i, s in enumerate(ref.astype(np.float32)):
        if s < best:
            best, keep = s, i
    c = int(chosen.item())
    assert c == keep or abs(ref[c] - ref[keep]) <= 4e-5 * abs(ref[keep]), (c, keep, ref[c], ref[keep])
    f = np.float32(1.0 - (c * 0.01))
    lo, hi = np.float32(x.min().item()) * f, np.float32(x.max().item()) * f
    beta = np.float32(0.9)
    omb = np.float32(1.0) - beta
    assert state[0].item() == np.float32(np.float32(0.0) * beta + lo * omb)
    assert state[1].item() == np.float32(np.float32(0.0) * beta + hi * omb)
    assert state[3].item() == np.float32(1.0) * beta


def test_module_follows_the_oracle_module_over_several_steps():
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    g = torch.Generator().manual_seed(5)
    ours, ref = qm.QuantAct_MSE(4).to(DEV), fq_torch.OracleQuantActMSE(4)
    for step in range(4):
        if step == 3:
            ours.fix()
            ref.fix()
        x = torch.relu(torch.randn(4, 8, 10, 10, generator=g) * (1.0 + step))
        y, y_ref = ours(x.to(DEV)), ref(x)
        for name in ("x_min", "x_max", "beta_t"):
            np.testing.assert_allclose(getattr(ours, name).cpu().numpy(), getattr(ref, name).numpy(), rtol=1e-6)
        np.testing.assert_allclose(y.cpu().numpy(), y_ref.numpy(), rtol=1e-5, atol=1e-6)
    assert ours.cur_x_min.dim() == 0 and ours.cur_x_max.dim() == 0


def test_search_rejects_bad_arguments():
    from ood_dfq_b200 import ops
    state = [t.to(DEV) for t in (torch.zeros(1), torch.zeros(1), torch.tensor([0.9]), torch.ones(1))]
    with pytest.raises(RuntimeError):
        ops.act_mse_search(torch.zeros(4), 4, *state)                      # CPU tensor
    with pytest.raises(RuntimeError):
        ops.act_mse_search(torch.zeros(4, device=DEV), 4, *state, steps=200)
    with pytest.raises(RuntimeError):
        ops.act_mse_search(torch.zeros(0, device=DEV), 4, *state)
