"""Batch assembly (crop -> bilinear resize -> grey-to-RGB repeat -> flip), CPU side.

* the oracle restatement (oracle/augment_torch.py) against the golden vectors torchvision itself produced
  (tools/make_golden.py::gen_augment) and, where torchvision is importable, against the live transform objects;
* the product's host code (ood_dfq_b200/augment.py): the draw-for-draw restatement of get_params + flip, the
  batched sampler, box validation;
* the product's per-pixel kernel body (csrc/augment_core.h) compiled for the host by g++ -- the same index
  arithmetic and blend the sm_100a kernel runs -- against the golden vectors in both layouts and all channel modes.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import augment_torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETS = ["rgb32", "grey28", "tall160", "rect", "big200"]


def case(golden, tag):
    z = golden("augment")
    return {k[len(tag) + 1:]: z[k] for k in z if k.startswith(tag + "_")}


# ----------------------------------------------------------------------------------------- oracle vs torchvision
@pytest.mark.parametrize("tag", SETS)
def test_oracle_reproduces_torchvision_outputs(golden, tag):
    """Same crop / interpolate / repeat / flip calls as torchvision: bit-exact with the recorded pipeline output."""
    c = case(golden, tag)
    images = torch.from_numpy(c["images"])
    size = c["out"].shape[2:]
    y = augment_torch.batch(images, c["index"], c["boxes"], c["flips"], size, channels=c["out"].shape[1])
    assert torch.equal(y, torch.from_numpy(c["out"]))
    y64 = augment_torch.batch(images, c["index"], c["boxes"], c["flips"], size, channels=c["out"].shape[1], dtype=torch.float64)
    assert torch.equal(y64.float(), torch.from_numpy(c["out_exact"]))


@pytest.mark.parametrize("tag", SETS)
def test_draws_reproduce_torchvision_streams(golden, tag):
    """From the recorded seed, the oracle's and the product's restatements of get_params + the flip coin yield the
    recorded boxes and flips (tall160: every attempt is rejected, so this is the central-crop fallback)."""
    from ood_dfq_b200 import augment
    c = case(golden, tag)
    h, w = c["images"].shape[2:]
    n = len(c["index"])
    torch.manual_seed(int(c["seed"]))
    drawn = [augment_torch.draw(h, w) for _ in range(n)]
    assert np.array_equal(np.array([d[0] for d in drawn], dtype=np.int32), c["boxes"])
    assert np.array_equal(np.array([d[1] for d in drawn], dtype=np.uint8), c["flips"])
    torch.manual_seed(int(c["seed"]))
    boxes, flips = augment.random_resized_crop_params(n, h, w)
    assert np.array_equal(boxes, c["boxes"]) and np.array_equal(flips, c["flips"])
    g = torch.Generator().manual_seed(int(c["seed"]))           # an explicit generator walks the same stream
    boxes, flips = augment.random_resized_crop_params(n, h, w, generator=g)
    assert np.array_equal(boxes, c["boxes"]) and np.array_equal(flips, c["flips"])


def test_draws_against_live_torchvision():
    tv = pytest.importorskip("torchvision.transforms")
    from ood_dfq_b200 import augment
    for (h, w), seed in (((32, 32), 3), ((224, 224), 4), ((28, 28), 5), ((17, 40), 6), ((64, 8), 7)):
        img = torch.zeros(1, h, w)
        torch.manual_seed(seed)
        want_b, want_f = [], []
        for _ in range(40):
            want_b.append(tv.RandomResizedCrop.get_params(img, [0.5, 1.0], [3 / 4, 4 / 3]))
            want_f.append(int(torch.rand(1) < 0.5))
        torch.manual_seed(seed)
        boxes, flips = augment.random_resized_crop_params(40, h, w)
        assert np.array_equal(boxes, np.array(want_b, dtype=np.int32)) and np.array_equal(flips, np.array(want_f, dtype=np.uint8))


def test_batched_sampler_has_the_same_distribution():
    from ood_dfq_b200 import augment
    g = torch.Generator().manual_seed(11)
    n, h, w = 20000, 224, 224
    b1, f1 = augment.random_resized_crop_params_batched(n, h, w, generator=g)
    augment.check_boxes(b1, h, w, (h, w))
    g2 = torch.Generator().manual_seed(12)
    b2, f2 = augment.random_resized_crop_params(4000, h, w, generator=g2)
    area1, area2 = b1[:, 2] * b1[:, 3] / (h * w), b2[:, 2] * b2[:, 3] / (h * w)
    assert abs(area1.mean() - area2.mean()) < 0.01 and abs(area1.std() - area2.std()) < 0.01
    assert area1.min() >= 0.49 and area1.max() <= 1.0
    r1, r2 = np.log(b1[:, 3] / b1[:, 2]), np.log(b2[:, 3] / b2[:, 2])
    assert abs(r1.mean() - r2.mean()) < 0.01 and abs(r1.std() - r2.std()) < 0.01
    assert abs(f1.mean() - 0.5) < 0.02 and abs(f2.mean() - 0.5) < 0.03
    # positions are uniform over the admissible range
    slack = (h - b1[:, 2]) > 20
    assert abs((b1[slack, 0] / (h - b1[slack, 2])).mean() - 0.5) < 0.02
    # an image no attempt fits: central crop, clamped to the ratio bounds
    b3, _ = augment.random_resized_crop_params_batched(5, 224, 24, generator=g)
    assert (b3 == np.array([96, 0, 32, 24])).all()


def test_box_validation():
    from ood_dfq_b200 import augment
    ok = np.array([[0, 0, 8, 8], [2, 3, 6, 5]], dtype=np.int32)
    augment.check_boxes(ok, 8, 8, (8, 8))
    for bad in ([[0, 0, 9, 8]], [[-1, 0, 4, 4]], [[5, 5, 4, 4]], [[0, 0, 0, 4]]):
        with pytest.raises(ValueError):
            augment.check_boxes(np.array(bad, dtype=np.int32), 8, 8, (8, 8))
    with pytest.raises(ValueError, match="antialias"):
        augment.check_boxes(ok, 8, 8, (4, 4))
    augment.check_boxes(ok, 8, 8, (4, 4), allow_downscale=True)


def test_device_shards_refuse_the_cpu():
    from ood_dfq_b200 import augment
    with pytest.raises(RuntimeError, match="no CPU path"):
        augment.DeviceShards(np.zeros((4, 3, 8, 8), np.float32), np.zeros(4, np.int64), 2, "cpu")


# ----------------------------------------------------------------------------------------- kernel body on the host
@pytest.fixture(scope="module")
def host_kernel(tmp_path_factory):
    """csrc/augment_core.h compiled by g++ (no FMA contraction, like the oracle's C build)."""
    out = tmp_path_factory.mktemp("augment_host") / "libaugment_host.so"
    src = os.path.join(ROOT, "tests", "host", "augment_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(out), src], check=True)
    lib = C.CDLL(str(out))
    vp = C.c_void_p
    lib.augment_host.argtypes = [vp, C.c_longlong, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp] + [C.c_int] * 7
    lib.augment_host.restype = C.c_int
    lib.augment_host_backward.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int, C.c_int, vp, vp, vp] + [C.c_int] * 6
    lib.augment_host_backward.restype = C.c_int

    def draws(index, boxes, flips):
        return (np.ascontiguousarray(index, np.int64), np.ascontiguousarray(boxes, np.int32),
                np.ascontiguousarray(flips, np.uint8))

    def run(images, index, boxes, flips, size, c_out, nhwc, px, src_nhwc=False):
        images = np.ascontiguousarray(images, np.float32)
        index, boxes, flips = draws(index, boxes, flips)
        m, c_in, h, w = images.shape
        stored = np.ascontiguousarray(images.transpose(0, 2, 3, 1)) if src_nhwc else images
        n = len(index)
        shape = (n, size[0], size[1], c_out) if nhwc else (n, c_out, size[0], size[1])
        out = np.full(shape, np.nan, np.float32)
        rc = lib.augment_host(stored.ctypes.data, m, c_in, h, w, index.ctypes.data, boxes.ctypes.data, flips.ctypes.data,
                              out.ctypes.data, n, c_out, size[0], size[1], int(nhwc), px, int(src_nhwc))
        assert rc == 0
        return out.transpose(0, 3, 1, 2) if nhwc else out

    def run_backward(grad_out, image_shape, index, boxes, flips, nhwc, src_nhwc=False):
        """grad_out [N,C_out,OH,OW] (logical NCHW) -> gradient w.r.t. the image set, logical [M,C,H,W]."""
        grad_out = np.asarray(grad_out, np.float32)
        index, boxes, flips = draws(index, boxes, flips)
        m, c_in, h, w = image_shape
        n, c_out, oh, ow = grad_out.shape
        go = np.ascontiguousarray(grad_out.transpose(0, 2, 3, 1)) if nhwc else np.ascontiguousarray(grad_out)
        gi = np.zeros((m, h, w, c_in) if src_nhwc else (m, c_in, h, w), np.float32)
        rc = lib.augment_host_backward(go.ctypes.data, gi.ctypes.data, m, c_in, h, w, index.ctypes.data, boxes.ctypes.data,
                                       flips.ctypes.data, n, c_out, oh, ow, int(nhwc), int(src_nhwc))
        assert rc == 0
        return gi.transpose(0, 3, 1, 2) if src_nhwc else gi
    run.backward = run_backward
    return run


def tolerances(c):
    """(tight, loose): against the double-precision result only the blend's fp32 roundings remain; against the fp32
    reference the source coordinate itself carries half an ulp at its magnitude (1.5e-5 spacing at 200)."""
    peak = float(np.abs(c["images"]).max())
    side = max(c["images"].shape[2:])
    return 4e-7 * peak, 2.0 ** -23 * side * 2 * peak + 4e-7 * peak


@pytest.mark.parametrize("tag", SETS)
@pytest.mark.parametrize("nhwc,px", [(True, 4), (True, 1), (False, 1), (False, 4)])
def test_kernel_body_matches_torchvision(golden, host_kernel, tag, nhwc, px):
    c = case(golden, tag)
    size, c_out = c["out"].shape[2:], c["out"].shape[1]
    y = host_kernel(c["images"], c["index"], c["boxes"], c["flips"], size, c_out, nhwc, px)
    tight, loose = tolerances(c)
    assert np.isfinite(y).all()
    assert np.abs(y - c["out_exact"]).max() <= tight
    assert np.abs(y - c["out"]).max() <= loose


def test_kernel_body_grey_stays_grey_or_repeats(golden, host_kernel):
    c = case(golden, "grey28")
    size = c["out"].shape[2:]
    y3 = host_kernel(c["images"], c["index"], c["boxes"], c["flips"], size, 3, True, 4)
    y1 = host_kernel(c["images"], c["index"], c["boxes"], c["flips"], size, 1, True, 4)
    assert np.array_equal(y3[:, 0], y3[:, 1]) and np.array_equal(y3[:, 0], y3[:, 2]) and np.array_equal(y1[:, 0], y3[:, 0])


def test_kernel_body_identity_and_mirror(host_kernel):
    """The whole image as the box is the identity resize (every weight exactly 0 or 1): bit-exact copy / mirror."""
    rng = np.random.default_rng(5)
    images = rng.standard_normal((3, 3, 9, 14)).astype(np.float32)
    index = np.array([2, 0, 1, 1], np.int64)
    boxes = np.tile(np.array([0, 0, 9, 14], np.int32), (4, 1))
    flips = np.array([0, 1, 0, 1], np.uint8)
    for nhwc, px in ((True, 4), (False, 1)):
        y = host_kernel(images, index, boxes, flips, (9, 14), 3, nhwc, px)
        for n in range(4):
            want = images[index[n]][..., ::-1] if flips[n] else images[index[n]]
            assert np.array_equal(y[n], want)


def test_kernel_body_ragged_totals_and_corrupt_entries(host_kernel):
    """Pixel counts that are not a multiple of the 4-pixel group, a one-pixel crop, and out-of-range boxes / indices
    folded into the image set instead of read out of bounds."""
    rng = np.random.default_rng(6)
    images = rng.standard_normal((2, 1, 5, 7)).astype(np.float32)
    index = np.array([1, 0, 7, -3], np.int64)                        # 7 -> image 1, -3 -> image 0
    boxes = np.array([[2, 3, 1, 1], [0, 0, 5, 7], [4, 6, 9, 9], [-2, -2, 3, 3]], np.int32)
    flips = np.array([0, 0, 1, 0], np.uint8)
    y = host_kernel(images, index, boxes, flips, (3, 3), 1, True, 4)                 # 4 * 9 = 36 pixels, groups of 4
    y1 = host_kernel(images, index, boxes, flips, (3, 3), 1, False, 1)
    assert np.isfinite(y).all() and np.array_equal(y, y1)
    assert np.all(y[0] == images[1, 0, 2, 3])                        # one-pixel crop: constant
    fixed = np.array([[2, 3, 1, 1], [0, 0, 5, 7], [0, 0, 5, 7], [0, 0, 3, 3]], np.int32)
    want = augment_torch.batch(torch.from_numpy(images), [1, 0, 1, 0], fixed, flips, (3, 3), channels=1,
                               dtype=torch.float64).float().numpy()
    # box 1 is a down-scaling (5x7 -> 3x3): the plain bilinear filter, not torchvision's antialiased one
    for n in (0, 3):
        np.testing.assert_allclose(y[n], want[n], atol=2e-6)
    plain = torch.nn.functional.interpolate(torch.from_numpy(images[1:2]).double(), size=(3, 3), mode="bilinear",
                                            align_corners=False).flip(-1).float().numpy()
    np.testing.assert_allclose(y[2], plain[0], atol=2e-6)
    y5 = host_kernel(images, index[:1], boxes[1:2], flips[:1], (5, 5), 1, True, 4)   # 25 pixels: ragged last group
    assert np.isfinite(y5).all()


@pytest.mark.parametrize("tag", SETS)
def test_kernel_body_reads_channels_last_image_sets(golden, host_kernel, tag):
    """The stored set may be channels_last (the distillation loop augments the optimised batch itself): same values."""
    c = case(golden, tag)
    size, c_out = c["out"].shape[2:], c["out"].shape[1]
    a = host_kernel(c["images"], c["index"], c["boxes"], c["flips"], size, c_out, True, 4)
    b = host_kernel(c["images"], c["index"], c["boxes"], c["flips"], size, c_out, True, 4, src_nhwc=True)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("tag", SETS)
@pytest.mark.parametrize("nhwc,src_nhwc", [(True, False), (False, False), (True, True)])
def test_kernel_body_backward_matches_torchvision_autograd(golden, host_kernel, tag, nhwc, src_nhwc):
    """Gradient scatter against what autograd computes through torchvision's own pipeline in double precision
    (tools/make_golden.py): taps shared by up to ~(out/in + 1)^2 outputs are summed in fp32 here."""
    c = case(golden, tag)
    # big200 stores ONE cotangent plane that the generator applied to each of the three (repeated) output channels
    cot = np.repeat(c["cotangent"], 3, axis=1) if tag == "big200" else c["cotangent"]
    g = host_kernel.backward(cot, c["images"].shape, c["index"], c["boxes"], c["flips"], nhwc, src_nhwc)
    scale = float(np.abs(c["grad_exact"]).max())
    assert np.abs(g - c["grad_exact"]).max() <= 2e-6 * scale


def test_kernel_body_backward_is_the_adjoint_of_the_forward(host_kernel):
    """<forward(x), g> == <x, backward(g)> for random x, g (all channel modes, flips, shared images)."""
    rng = np.random.default_rng(12)
    for c_in, c_out in ((3, 3), (1, 3), (1, 1)):
        images = rng.standard_normal((3, c_in, 11, 13)).astype(np.float32)
        index = np.array([2, 2, 0, 1, 2], np.int64)                 # image 2 feeds three samples: gradients add up
        boxes = np.array([[0, 0, 11, 13], [2, 3, 6, 7], [1, 0, 9, 13], [4, 5, 5, 5], [0, 6, 11, 7]], np.int32)
        flips = np.array([0, 1, 1, 0, 1], np.uint8)
        y = host_kernel(images, index, boxes, flips, (11, 13), c_out, True, 4)
        g = rng.standard_normal(y.shape).astype(np.float32)
        gx = host_kernel.backward(g, images.shape, index, boxes, flips, True)
        lhs, rhs = float((y.astype(np.float64) * g).sum()), float((images.astype(np.float64) * gx).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_oracle_gradient_reproduces_torchvision_autograd(golden):
    """The oracle restatement is differentiable the same way (ATen's interpolate backward): its double-precision
    gradient equals the recorded one."""
    c = case(golden, "rect")
    leaf = torch.from_numpy(c["images"]).double().requires_grad_(True)
    y = augment_torch.batch(leaf, c["index"], c["boxes"], c["flips"], c["out"].shape[2:], channels=3)
    (y * torch.from_numpy(c["cotangent"]).double()).sum().backward()
    assert torch.equal(leaf.grad.float(), torch.from_numpy(c["grad_exact"]))


def test_kernel_body_random_geometries(host_kernel):
    """hypothesis: any image / box / output geometry (up- AND down-scaling, one-pixel sides, flips, every channel mode
    and layout) against the plain bilinear filter (``F.interpolate(..., antialias=False)``) evaluated in double."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(deadline=None, max_examples=150, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(seed=st.integers(0, 2 ** 31), h=st.integers(1, 33), w=st.integers(1, 33), oh=st.integers(1, 40),
           ow=st.integers(1, 40), mode=st.sampled_from([(3, 3), (1, 3), (1, 1)]), nhwc=st.booleans(), src_nhwc=st.booleans())
    def check(seed, h, w, oh, ow, mode, nhwc, src_nhwc):
        rng = np.random.default_rng(seed)
        c_in, c_out = mode
        m, n = 3, 5
        images = rng.standard_normal((m, c_in, h, w)).astype(np.float32)
        index = rng.integers(0, m, n).astype(np.int64)
        bh, bw = rng.integers(1, h + 1, n), rng.integers(1, w + 1, n)
        top, left = rng.integers(0, h - bh + 1), rng.integers(0, w - bw + 1)
        boxes = np.stack([top, left, bh, bw], 1).astype(np.int32)
        flips = rng.integers(0, 2, n).astype(np.uint8)
        y = host_kernel(images, index, boxes, flips, (oh, ow), c_out, nhwc, 4 if nhwc else 1, src_nhwc=src_nhwc)
        want = torch.stack([augment_torch.resized_crop_flip(torch.from_numpy(images[index[j]]).double(), boxes[j], bool(flips[j]),
                                                            (oh, ow), antialias=False)[:c_out] for j in range(n)]).float().numpy()
        assert np.abs(y - want).max() <= 1e-6 * max(1.0, float(np.abs(images).max()))
        # and the gradient scatter is the exact adjoint of that map
        g = rng.standard_normal(y.shape).astype(np.float32)
        gx = host_kernel.backward(g, images.shape, index, boxes, flips, nhwc, src_nhwc)
        lhs, rhs = float((y.astype(np.float64) * g).sum()), float((images.astype(np.float64) * gx).sum())
        assert abs(lhs - rhs) <= 2e-5 * max(1.0, float(np.abs(y.astype(np.float64) * g).sum()))
    check()
