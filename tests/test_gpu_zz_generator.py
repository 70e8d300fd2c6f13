"""Warm-up phase iteration (step.GeneratorStep, trainer_direct.py:459-488) on the GPU mirror against the CPU oracle:
generator loss with the fused BN-statistics loss (trainer flavour) and the student's calibrating QuantAct path.

Written in a session without GPU minutes (file name: runs last); the host code is pinned bit for bit to the
reference's own source on CPU (tests/test_dropin_reference_code.py)."""
import copy

import pytest
import torch
from torch import nn

from oracle import bns_torch, fq_torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


class TinyGenerator(nn.Module):
    """Stand-in with the reference generator's interface ``G(z, labels)`` (main_direct.py:52-88 is not importable)."""

    def __init__(self, n_classes=10, latent=32, side=32):
        super().__init__()
        self.emb = nn.Embedding(n_classes, latent)
        self.side = side
        self.fc = nn.Linear(latent, 16 * (side // 4) ** 2)
        self.body = nn.Sequential(nn.BatchNorm2d(16), nn.Upsample(scale_factor=4), nn.Conv2d(16, 3, 3, padding=1), nn.Tanh())

    def forward(self, z, labels):
        h = self.fc(self.emb(labels) * z).view(z.shape[0], 16, self.side // 4, self.side // 4)
        return self.body(h)


def test_generator_phase_matches_the_cpu_oracle():
    from ood_dfq_b200 import bns, nets, step, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    gen = TinyGenerator()
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    s_cpu = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=fq_torch)
    s_gpu = surgery.quantize_model(copy.deepcopy(teacher), 4, 4).to(DEV)
    t_cpu, t_gpu = copy.deepcopy(teacher), copy.deepcopy(teacher).to(DEV)
    g_cpu, g_gpu = copy.deepcopy(gen), copy.deepcopy(gen).to(DEV)
    ref = step.GeneratorStep(g_cpu, t_cpu, s_cpu, bns_torch.StatTap(t_cpu), latent_dim=32, n_classes=10)
    ours = step.GeneratorStep(g_gpu, t_gpu, s_gpu, bns.BNStatLoss(t_gpu), latent_dim=32, n_classes=10)
    for it in range(3):
        torch.manual_seed(50 + it)                    # z and labels are drawn on the host: same batch on both sides
        want = ref()
        torch.manual_seed(50 + it)
        got = ours()
        for a, b in zip(got, want):
            assert abs(a.item() - b.item()) <= 2e-3 * abs(b.item()) + 1e-6, (it, a.item(), b.item())
    acts_c = [m for m in s_cpu.modules() if type(m) is fq_torch.OracleQuantAct]
    acts_g = [m for m in s_gpu.modules() if type(m) is qm.QuantAct]
    assert len(acts_c) == len(acts_g) > 10
    for c, g in zip(acts_c, acts_g):
        assert torch.equal(c.beta_t, g.beta_t.cpu())                 # three calibrating forwards on both sides
        assert abs(g.x_max.item() - c.x_max.item()) <= 2e-2 * abs(c.x_max.item()) + 1e-4
        assert abs(g.x_min.item() - c.x_min.item()) <= 2e-2 * abs(c.x_max.item()) + 1e-4
