"""Batch assembly on the device (csrc/augment.cu) against torchvision's recorded outputs and the CPU oracle.

Runs last among the GPU tests on purpose (file name): the kernel was written in a session that had no GPU minutes
left; its per-pixel body is verified on the host (tests/test_augment_cpu.py), this file adds the launch geometry.
"""
import numpy as np
import pytest
import torch

from oracle import augment_torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
SETS = ["rgb32", "grey28", "tall160", "rect", "big200"]


def case(golden, tag):
    z = golden("augment")
    return {k[len(tag) + 1:]: z[k] for k in z if k.startswith(tag + "_")}


def run(ops, images, index, boxes, flips, size, channels, channels_last):
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a)).to(t).to(DEV)
    return ops.crop_resize_flip(d(images, torch.float32), d(index, torch.int64), d(boxes, torch.int32),
                                d(flips, torch.uint8), size, channels=channels, channels_last=channels_last)


@pytest.mark.parametrize("tag", SETS)
@pytest.mark.parametrize("channels_last", [True, False])
def test_golden_augment(golden, tag, channels_last):
    """Tolerances as in tests/test_augment_cpu.py: 4e-7 * peak against torchvision in double precision (only the
    blend's fp32 roundings remain), coordinate spacing against torchvision in fp32."""
    from ood_dfq_b200 import ops
    c = case(golden, tag)
    size, c_out = tuple(c["out"].shape[2:]), c["out"].shape[1]
    y = run(ops, c["images"], c["index"], c["boxes"], c["flips"], size, c_out, channels_last)
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    assert y.shape == c["out"].shape and y.is_contiguous(memory_format=fmt)
    y = y.cpu().numpy()
    peak, side = float(np.abs(c["images"]).max()), max(c["images"].shape[2:])
    assert np.abs(y - c["out_exact"]).max() <= 4e-7 * peak
    assert np.abs(y - c["out"]).max() <= 2.0 ** -23 * side * 2 * peak + 4e-7 * peak


@pytest.mark.parametrize("shape,batch", [((40, 3, 32, 32), 256), ((12, 1, 28, 28), 64), ((6, 3, 224, 224), 16),
                                         ((5, 3, 33, 47), 7)])
def test_seeded_batches_vs_oracle(shape, batch):
    """Seeded image sets at the configs' shapes, boxes and flips from the product's own sampler, against the oracle
    evaluated in double precision; identity boxes reproduce the stored image bit for bit."""
    from ood_dfq_b200 import augment, ops
    g = torch.Generator().manual_seed(sum(shape))
    images = torch.randn(shape, generator=g)
    m, c, h, w = shape
    index = torch.randint(0, m, (batch,), generator=g)
    boxes, flips = augment.random_resized_crop_params_batched(batch, h, w, generator=g)
    want = augment_torch.batch(images, index.tolist(), boxes, flips, (h, w), channels=3, dtype=torch.float64).float()
    for channels_last in (True, False):
        y = run(ops, images.numpy(), index.numpy(), boxes, flips, (h, w), 3, channels_last)
        assert (y.cpu() - want).abs().max().item() <= 4e-7 * images.abs().max().item()
    ident = np.tile(np.array([0, 0, h, w], np.int32), (batch, 1))
    y = run(ops, images.numpy(), index.numpy(), ident, np.zeros(batch, np.uint8), (h, w), 3, True).cpu()
    src = images[index]
    assert torch.equal(y, src.expand(-1, 3, -1, -1) if c == 1 else src)
    y = run(ops, images.numpy(), index.numpy(), ident, np.ones(batch, np.uint8), (h, w), 3, True).cpu()
    assert torch.equal(y, (src.expand(-1, 3, -1, -1) if c == 1 else src).flip(-1))


def test_argument_errors():
    from ood_dfq_b200 import ops
    img = torch.zeros(2, 3, 8, 8, device=DEV)
    idx = torch.zeros(2, dtype=torch.int64, device=DEV)
    box = torch.tensor([[0, 0, 8, 8]] * 2, dtype=torch.int32, device=DEV)
    flip = torch.zeros(2, dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.crop_resize_flip(img.cpu(), idx, box, flip, 8)                      # no CPU path
    with pytest.raises(RuntimeError):
        ops.crop_resize_flip(img, idx.int(), box, flip, 8)                      # index must be int64
    with pytest.raises(RuntimeError):
        ops.crop_resize_flip(img, idx, box[:1], flip, 8)                        # one box per sample
    with pytest.raises(RuntimeError, match="channels"):
        ops.crop_resize_flip(img, idx, box, flip, 8, channels=1)                # 3 -> 1 is not a mode
    assert ops.crop_resize_flip(img, idx[:0], box[:0], flip[:0], 8).shape == (0, 3, 8, 8)


def test_device_shards_stream_batches():
    """DeviceShards: every sample of the rank's split appears once per epoch with its label, batches keep the
    requested memory format, and without augmentation the batch is the plain gather."""
    from ood_dfq_b200 import augment, shards
    rng = np.random.default_rng(3)
    images = rng.standard_normal((50, 1, 28, 28)).astype(np.float32)
    labels = np.arange(50, dtype=np.int64)
    ds = augment.DeviceShards(images, labels, batch=8, device=DEV, rank=1, world=2, seed=5, augment=False)
    want_idx = shards.rank_indices(50, 1, 2, 0, True, 5)
    seen = []
    for x, y in ds:
        assert x.shape == (8, 3, 28, 28) and x.is_contiguous(memory_format=torch.channels_last)
        idx = y.cpu().numpy()
        seen.extend(idx.tolist())
        assert torch.equal(x.cpu(), torch.from_numpy(images[idx]).expand(-1, 3, -1, -1))
    assert seen == want_idx[:len(seen)].tolist() and len(seen) == len(ds) * 8 == 24
    aug = augment.DeviceShards(images, labels, batch=8, device=DEV, seed=5)
    x, y = next(iter(aug))
    assert x.shape == (8, 3, 28, 28) and torch.isfinite(x).all()
    lo, hi = images[y.cpu().numpy()].min(), images[y.cpu().numpy()].max()
    assert x.min().item() >= lo - 1e-6 and x.max().item() <= hi + 1e-6       # bilinear blends stay inside the data range


@pytest.mark.parametrize("tag", SETS)
@pytest.mark.parametrize("grad_cl,src_cl", [(True, False), (False, False), (True, True)])
def test_golden_augment_backward(golden, tag, grad_cl, src_cl):
    """Gradient scatter against autograd through torchvision's own pipeline in double precision (fp32 atomics here,
    any order: 2e-6 of the largest entry)."""
    from ood_dfq_b200 import ops
    c = case(golden, tag)
    cot = np.repeat(c["cotangent"], 3, axis=1) if tag == "big200" else c["cotangent"]
    go = torch.from_numpy(np.ascontiguousarray(cot)).to(DEV)
    go = go.contiguous(memory_format=torch.channels_last) if grad_cl else go
    like = torch.empty(c["images"].shape, device=DEV)
    like = like.contiguous(memory_format=torch.channels_last) if src_cl else like
    d = lambda a, t: torch.from_numpy(np.ascontiguousarray(a)).to(t).to(DEV)
    g = ops.crop_resize_flip_backward(go, like, d(c["index"], torch.int64), d(c["boxes"], torch.int32),
                                      d(c["flips"], torch.uint8))
    assert g.shape == like.shape and g.stride() == like.stride()
    scale = float(np.abs(c["grad_exact"]).max())
    assert np.abs(g.cpu().numpy() - c["grad_exact"]).max() <= 2e-6 * scale
    # accumulation into an existing gradient
    base = torch.randn_like(like)
    g2 = ops.crop_resize_flip_backward(go, like, d(c["index"], torch.int64), d(c["boxes"], torch.int32),
                                       d(c["flips"], torch.uint8), accumulate_into=base.clone())
    assert (g2 - base - g).abs().max().item() <= 4e-6 * scale + 1e-6


@pytest.mark.parametrize("src_cl", [False, True])
def test_differentiable_augmentation_matches_oracle_autograd(src_cl):
    """augment.crop_resize_flip under autograd (the distillation loop's RHF(RRC(x[j])) on the optimised batch itself,
    data_generate/distill_data.py:197-227): output and the gradient reaching the images against the oracle evaluated in
    double precision."""
    from ood_dfq_b200 import augment
    g = torch.Generator().manual_seed(31)
    b, h, w = 6, 40, 40
    x = torch.randn(b, 3, h, w, generator=g) / 5
    boxes, flips = augment.random_resized_crop_params(b, h, w, scale=(0.4, 1.0), generator=g)
    index = torch.arange(b)
    cot = torch.randn(b, 3, h, w, generator=g)
    leaf = x.double().requires_grad_(True)
    y_ref = augment_torch.batch(leaf, index.tolist(), boxes, flips, (h, w), channels=3)
    (y_ref * cot.double()).sum().backward()
    xd = x.to(DEV)
    xd = (xd.contiguous(memory_format=torch.channels_last) if src_cl else xd).requires_grad_(True)
    y = augment.crop_resize_flip(xd, index.to(DEV), torch.from_numpy(boxes).to(DEV), torch.from_numpy(flips).to(DEV))
    (y * cot.to(DEV)).sum().backward()
    assert (y.detach().cpu() - y_ref.detach().float()).abs().max().item() <= 4e-7 * x.abs().max().item()
    assert xd.grad.shape == xd.shape and xd.grad.stride() == xd.stride()
    scale = leaf.grad.abs().max().item()
    assert (xd.grad.cpu() - leaf.grad.float()).abs().max().item() <= 2e-6 * scale


def test_augmented_distillation_iteration_matches_the_cpu_arm():
    """step.DistillStep(augment=augment.batch_augmenter()) on the GPU against the same iteration with the oracle's
    torch augmentation on CPU: equal seeds give equal crops, losses agree, gradients point the same way."""
    import copy
    import random

    from ood_dfq_b200 import augment, bns, nets, step
    from oracle import bns_torch
    torch.manual_seed(4)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 3, 32, 32, generator=g) / 5
    labels = torch.randint(0, 10, (8,), generator=g)
    torch.backends.cudnn.allow_tf32 = False

    def oracle_augment(t, boxes, flips):
        return augment_torch.batch(t, range(t.shape[0]), boxes, flips, t.shape[2:], channels=3)
    cpu = copy.deepcopy(teacher)
    ref = step.DistillStep(cpu, bns_torch.StatTap(cpu), x, labels, augment=oracle_augment, augment_p=1.0)
    gpu = copy.deepcopy(teacher).to(DEV)
    ours = step.DistillStep(gpu, bns.BNStatLoss(gpu), x.to(DEV), labels.to(DEV), augment=augment.batch_augmenter(),
                            augment_p=1.0)
    random.seed(1)
    torch.manual_seed(2)
    l_ref = ref().item()
    random.seed(1)
    torch.manual_seed(2)
    l_ours = ours().item()
    assert abs(l_ours - l_ref) <= 1e-3 * abs(l_ref), (l_ours, l_ref)
    cos = torch.nn.functional.cosine_similarity(ref.images.grad.flatten(), ours.images.grad.cpu().flatten(), dim=0).item()
    assert cos > 0.999, cos
    with pytest.raises(ValueError, match="cannot be captured"):
        step.DistillStep(gpu, bns.BNStatLoss(gpu), x.to(DEV), labels.to(DEV), augment=augment.batch_augmenter(),
                         capturable=True)
