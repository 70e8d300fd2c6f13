"""Whole-plane average pool (csrc/gap.cu, ``fusion.fuse_global_avgpool``) against ``nn.AvgPool2d``: the kernel adds
the plane in window order in fp32 and divides once, like ATen's channels_last kernel, so forward and backward are
compared bit for bit; ``AdaptiveAvgPool2d(1)`` (ATen: tree-ordered mean) within fp32 rounding."""
import copy

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [(256, 512, 7, 7), (64, 64, 8, 8), (3, 8, 5, 3), (2, 12, 1, 2), (5, 1024, 2, 2), (1, 4, 14, 14), (130, 36, 4, 4)]


def bits(t):
    return t.detach().cpu().numpy().view(np.int32)


@pytest.mark.parametrize("shape", SHAPES)
def test_global_avgpool_is_bit_identical_to_avgpool2d(shape):
    from ood_dfq_b200 import fusion, ops
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3.0)
    x.view(-1)[::17] = 0.0
    x.view(-1)[5::29] *= -1e-30                     # tiny and negative values: the -0 / underflow corners
    x = x.to(DEV).contiguous(memory_format=torch.channels_last)
    pool = nn.AvgPool2d(kernel_size=shape[2:], stride=1)
    fused = copy.deepcopy(pool)
    assert fusion.fuse_global_avgpool(nn.Sequential(fused)) == 1 and isinstance(fused, nn.AvgPool2d)
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    n0 = ops.N.load().oodfq_launch_count()
    ya, yb = pool(a), fused(b)
    assert ops.N.load().oodfq_launch_count() == n0 + 1                 # the library took the call
    assert ya.shape == yb.shape == (shape[0], shape[1], 1, 1)
    assert np.array_equal(bits(ya), bits(yb))
    go = torch.randn(ya.shape, generator=g).to(DEV)
    go.view(-1)[::7] = -0.0
    go.view(-1)[3::11] = 1e-44                       # quotient underflows to +-0
    go.view(-1)[4::11] = -1e-44
    ya.backward(go)
    yb.backward(go)
    assert b.grad.is_contiguous(memory_format=torch.channels_last)
    assert np.array_equal(bits(a.grad), bits(b.grad))


def test_other_windows_and_layouts_keep_the_aten_path():
    from ood_dfq_b200 import fusion, ops
    x = torch.randn(4, 16, 8, 8, device=DEV)
    m = nn.Sequential(nn.AvgPool2d(4), nn.AvgPool2d(8, padding=0), nn.AdaptiveAvgPool2d(2), nn.AdaptiveAvgPool2d(1))
    assert fusion.fuse_global_avgpool(m) == 3 and fusion.fuse_global_avgpool(m) == 0
    assert type(m[2]) is nn.AdaptiveAvgPool2d
    n0 = ops.N.load().oodfq_launch_count()
    xl = x.contiguous(memory_format=torch.channels_last)
    assert torch.equal(m[0](xl), nn.functional.avg_pool2d(xl, 4))            # 4x4 window on an 8x8 plane: ATen
    assert torch.equal(m[1](x), nn.functional.avg_pool2d(x, 8))              # NCHW input: ATen
    assert ops.N.load().oodfq_launch_count() == n0
    ref = nn.functional.adaptive_avg_pool2d(xl, 1)
    got = m[3](xl)
    assert ops.N.load().oodfq_launch_count() == n0 + 1
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_carrier_network_with_the_fused_pool_gives_the_same_bits():
    from ood_dfq_b200 import fusion, nets
    torch.manual_seed(1)
    net = nets.resnet20_cifar(num_classes=10).to(DEV).to(memory_format=torch.channels_last).eval()
    fused = copy.deepcopy(net)
    assert fusion.fuse_global_avgpool(fused) == 1
    x = torch.randn(8, 3, 32, 32, device=DEV).contiguous(memory_format=torch.channels_last)
    a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    torch.backends.cudnn.deterministic = True
    try:
        ya, yb = net(a), fused(b)
        assert torch.equal(ya, yb)
        ya.square().sum().backward()
        yb.square().sum().backward()
        assert torch.equal(a.grad, b.grad)
        for (n1, p1), (_, p2) in zip(net.named_parameters(), fused.named_parameters()):
            assert torch.equal(p1.grad, p2.grad), n1
    finally:
        torch.backends.cudnn.deterministic = False
