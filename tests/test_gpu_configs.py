"""BASELINE.json configs as parity cases: every quantisation / BN site of the real networks.

cuDNN and the CPU convolution do not round identically, so a quantised network cannot be compared end to end
bit for bit (one ulp in a conv output can move a code across a rounding boundary).  The parity proper is
therefore teacher-forced: the CPU oracle model runs the config, every QuantAct input, BN input and weight it
saw is replayed through the CUDA path, and THOSE results must be bit-exact (codes, values, range state) or
within 1e-5 (statistics, loss, gradients).  An end-to-end run is then checked statistically.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import bits
from oracle import bns_torch, fq_torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def same_bits(t, ref):
    return np.array_equal(bits(t.detach().cpu().numpy()), bits(ref.detach().cpu().numpy()))


def build(net_name, classes, k, seed=1):
    from ood_dfq_b200 import nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.manual_seed(seed)
    base = nets.resnet18_small(3, classes) if net_name == "resnet18_small" else getattr(nets, net_name)(num_classes=classes)
    nets.perturb_bn_stats(base)
    cpu = surgery.quantize_model(copy.deepcopy(base), k, k, namespace=fq_torch).eval()
    gpu = surgery.quantize_model(copy.deepcopy(base), k, k, namespace=qm).to(DEV).eval()
    return cpu, gpu, qm


class SiteRecorder:
    """Inputs/outputs/state of every oracle QuantAct and inputs of every BN, in execution order."""

    def __init__(self, model):
        self.acts, self.bns = [], []
        for m in model.modules():
            if isinstance(m, fq_torch.OracleQuantAct):
                m.register_forward_hook(self._act)
            elif isinstance(m, torch.nn.BatchNorm2d):
                m.register_forward_hook(self._bn)

    def _act(self, m, inputs, output):
        self.acts.append((inputs[0].detach().clone(), output.detach().clone(),
                          torch.cat([m.x_min, m.x_max, m.beta_t]).clone()))

    def _bn(self, m, inputs, output):
        self.bns.append((inputs[0].detach().clone(), m.running_mean.clone(), m.running_var.clone()))

    def clear(self):
        self.acts.clear()
        self.bns.clear()


CONFIGS = [
    pytest.param("resnet20_cifar", 10, 4, (256, 3, 32, 32), 19, 21, id="cfg1-cifar10-resnet20-w4a4"),
    pytest.param("resnet20_cifar", 100, 4, (256, 3, 32, 32), 19, 21, id="cfg2-cifar100-resnet20-w4a4"),
    pytest.param("resnet18_small", 9, 2, (64, 3, 28, 28), 17, 20, id="cfg3-pathmnist-resnet18-w2a2"),
    pytest.param("resnet18_imagenet", 1000, 4, (16, 3, 224, 224), 17, 20, id="cfg4-imagenet-resnet18-w4a4-sample"),
]


@pytest.mark.parametrize("net_name,classes,k,shape,n_act,n_bn", CONFIGS)
def test_config_sites_teacher_forced(net_name, classes, k, shape, n_act, n_bn):
    from ood_dfq_b200 import bns, ops
    cpu, gpu, qm = build(net_name, classes, k)
    rec = SiteRecorder(cpu)
    gpu_acts = [m for m in gpu.modules() if type(m) is qm.QuantAct]
    cpu_acts = [m for m in cpu.modules() if type(m) is fq_torch.OracleQuantAct]
    assert len(gpu_acts) == n_act == len(cpu_acts)
    g = torch.Generator().manual_seed(0)
    # three calibrating steps then a frozen one
    for step in range(4):
        if step == 3:
            for m in cpu_acts + gpu_acts:
                m.fix()
        x = torch.randn(shape, generator=g)
        rec.clear()
        with torch.no_grad():
            cpu(x)
        assert len(rec.acts) == n_act and len(rec.bns) == n_bn
        # module order is not execution order in general: pair recorded sites with modules by a probe forward
        for (xin, yout, state), m_cpu in zip(rec.acts, _execution_order(cpu, cpu_acts, x)):
            m_gpu = gpu_acts[cpu_acts.index(m_cpu)]
            y = m_gpu(xin.to(DEV))
            assert same_bits(y, yout), (step, cpu_acts.index(m_cpu))
            st = torch.cat([m_gpu.x_min, m_gpu.x_max, m_gpu.beta_t])
            assert same_bits(st, state), (step, cpu_acts.index(m_cpu))
    # weights: every Quant_Conv2d / Quant_Linear through the multi-tensor launch AND the module cache
    cpu_w = [m for m in cpu.modules() if isinstance(m, (fq_torch.OracleQuantConv2d, fq_torch.OracleQuantLinear))]
    gpu_w = [m for m in gpu.modules() if isinstance(m, (qm.Quant_Conv2d, qm.Quant_Linear))]
    assert len(cpu_w) == len(gpu_w) == n_bn + 1
    for mc, mg in zip(cpu_w, gpu_w):
        assert same_bits(mg.quantized_weight(), mc.quantized_weight())
    res = ops.weight_fq_multi([m.weight for m in gpu_w], [k] * len(gpu_w), [False] * len(gpu_w), want_codes=True)
    for mc, r in zip(cpu_w, res):
        lo, hi = fq_torch.row_minmax(mc.weight)
        assert np.array_equal(r["codes"].cpu().numpy().astype(np.float32), fq_torch.codes(mc.weight.detach(), k, lo, hi).numpy())
    # BN sites of the last (frozen) forward: statistics, packed loss and the local input gradient
    sums, offs, counts, rms, rvs, xs = [], [0], [], [], [], []
    for xin, rm, rv in rec.bns:
        xg = xin.to(DEV)
        sums.append(ops.bn_stats_forward(xg, rm.to(DEV)))
        offs.append(offs[-1] + xin.shape[1])
        counts.append(float(xin.numel() // xin.shape[1]))
        rms.append(rm)
        rvs.append(rv)
        xs.append(xg)
    loss3, mean, var, gmean, gvar = ops.bns_loss(torch.cat(sums), torch.cat(rms).to(DEV), torch.cat(rms).to(DEV),
                                                 torch.cat(rvs).to(DEV), offs, counts)
    means, vars_ = zip(*[bns_torch.channel_stats(xin) for xin, _, _ in rec.bns])
    ref_loss = bns_torch.bns_loss_trainer(means, vars_, rms, rvs)
    np.testing.assert_allclose(loss3[0].item(), ref_loss.item(), rtol=1e-5)
    np.testing.assert_allclose(mean.cpu().numpy(), torch.cat(means).numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var.cpu().numpy(), torch.cat(vars_).numpy(), rtol=2e-5, atol=1e-7)
    L = len(rec.bns)
    for i in (0, L // 2, L - 1):                       # local gradient of three layers against the closed form
        xin, rm, rv = rec.bns[i]
        sl = slice(offs[i], offs[i + 1])
        gx = ops.bn_stats_backward(xs[i], None, mean[sl], gmean[sl], gvar[sl], counts[i])
        ref = bns_torch.bns_input_grad(xin, rm, rv, upstream=1.0 / L)
        scale = ref.abs().max().item()
        np.testing.assert_allclose(gx.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-5 * scale)


def _execution_order(model, acts, x):
    order = []
    hooks = [m.register_forward_hook(lambda mod, i, o: order.append(mod)) for m in acts]
    was = [m.running_stat for m in acts]
    for m in acts:                      # do not disturb the calibrated state while probing the order
        m.running_stat = False
    with torch.no_grad():
        model(x[:1])
    for m, w in zip(acts, was):
        m.running_stat = w
    for h in hooks:
        h.remove()
    return order


@pytest.mark.parametrize("net_name,classes,k,shape", [
    pytest.param("resnet20_cifar", 10, 4, (64, 3, 32, 32), id="cfg1-end-to-end"),
    pytest.param("resnet20_cifar", 100, 4, (64, 3, 32, 32), id="cfg2-end-to-end"),
    pytest.param("resnet18_small", 9, 2, (32, 3, 28, 28), id="cfg3-end-to-end"),
])
def test_config_end_to_end_statistical(net_name, classes, k, shape):
    """Free-running GPU model vs CPU oracle model: ranges agree to conv rounding, logits stay close."""
    cpu, gpu, qm = build(net_name, classes, k)
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(3)
    for step in range(3):
        x = torch.randn(shape, generator=g)
        with torch.no_grad():
            yc, yg = cpu(x), gpu(x.to(DEV))
    ca = [m for m in cpu.modules() if type(m) is fq_torch.OracleQuantAct]
    ga = [m for m in gpu.modules() if type(m) is qm.QuantAct]
    rc = torch.stack([torch.cat([m.x_min, m.x_max]) for m in ca])
    rg = torch.stack([torch.cat([m.x_min, m.x_max]) for m in ga]).cpu()
    np.testing.assert_allclose(rg.numpy(), rc.numpy(), rtol=2e-2, atol=1e-4)
    spread = yc.std().item()
    assert (yg.cpu() - yc).abs().mean().item() < 0.25 * spread


def test_config2_qat_step_against_the_cpu_oracle_step():
    """BASELINE configs[1] (cifar100_resnet20.hocon): one KD-style QAT iteration (trainer_direct.py:490-518 -- teacher
    forward, student forward, loss_fn_kd T=20 alpha=20 + feature alignment, sign perturbation, second pass, backward,
    SGD nesterov lr 1e-5) on the CUDA mirror against the same host code over the CPU oracle modules.  cuDNN and the
    CPU convolution round differently, so the comparison is statistical: loss within 2 %, the update points the same
    way."""
    from ood_dfq_b200 import nets, step, surgery
    cpu, gpu, qm = build("resnet20_cifar", 100, 4)
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(1)
    teacher = nets.perturb_bn_stats(nets.resnet20_cifar(num_classes=100))
    g = torch.Generator().manual_seed(7)
    calib = torch.randn(64, 3, 32, 32, generator=g)
    with torch.no_grad():
        for _ in range(3):
            cpu(calib)
            gpu(calib.to(DEV))
    surgery.freeze_model(cpu, fq_torch)
    surgery.freeze_model(gpu, qm)
    kw = dict(lr=1e-5, momentum=0.9, weight_decay=1e-4, temperature=20.0, alpha=20.0, lam=1000.0, eps=0.01,
              unit_types=(nets.ResUnit,))
    ref = step.QATStep(cpu, copy.deepcopy(teacher), **kw)
    ours = step.QATStep(gpu, copy.deepcopy(teacher).to(DEV), **kw)
    x = torch.randn(64, 3, 32, 32, generator=g)
    l_ref, l_ours = ref(x).item(), ours(x.to(DEV)).item()
    assert abs(l_ours - l_ref) <= 2e-2 * abs(l_ref), (l_ours, l_ref)
    cos = torch.nn.functional.cosine_similarity(ref.grads.flat, ours.grads.flat.cpu(), dim=0).item()
    assert cos > 0.9, cos


def test_config5_distillation_iteration():
    """BN-statistics distillation step (distill_data.py:229-275) with a quantize_model-wrapped teacher."""
    from ood_dfq_b200 import bns, step
    cpu, gpu, qm = build("resnet18_imagenet", 1000, 4)
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    calib = torch.randn(2, 3, 224, 224, generator=g) / 5
    with torch.no_grad():
        for _ in range(3):
            cpu(calib)
            gpu(calib.to(DEV))
    for m in cpu.modules():
        if type(m) is fq_torch.OracleQuantAct:
            m.fix()
    for m in gpu.modules():
        if type(m) is qm.QuantAct:
            m.fix()
    x = torch.randn(2, 3, 224, 224, generator=g) / 5            # distill_data.py:181
    labels = torch.randint(0, 1000, (2,), generator=g)
    ref = step.DistillStep(cpu, bns_torch.StatTap(cpu), x, labels)
    ours = step.DistillStep(gpu, bns.BNStatLoss(gpu), x.to(DEV), labels.to(DEV))
    l_ref, l_ours = ref().item(), ours().item()
    assert abs(l_ours - l_ref) <= 2e-2 * abs(l_ref), (l_ours, l_ref)
    # the gradient that reached the images (through cuDNN dgrad and every identity STE) points the same way
    cos = torch.nn.functional.cosine_similarity(ref.images.grad.flatten(), ours.images.grad.cpu().flatten(), dim=0).item()
    assert cos > 0.9, cos
