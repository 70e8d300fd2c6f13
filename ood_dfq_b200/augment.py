"""Device-side batch assembly: the data format and caller on the input side of the QAT step.

The reference keeps the synthetic images in host memory, augments them one sample at a time on DataLoader workers
(``direct_dataset.__getitem__``, ``main_direct.py:200-204``, with the torchvision pipeline of ``:158-169``), collates
and copies every batch to the GPU (``main_direct.py:525-533``).  Here the concatenated shards
(``shards.load_shards``) are uploaded once -- 180 GB of HBM hold ~300 000 fp32 224x224 images -- and a per-rank batch
is ONE kernel (``ops.crop_resize_flip``, ``csrc/augment.cu``) that gathers, crops, resizes, repeats grey to RGB, flips
and writes the channels_last batch the step consumes.  Per step the host ships only the sample indices, the crop
boxes and the flip bits (a few KB).

What stays on the host is the random draw.  ``random_resized_crop_params`` restates
``torchvision.transforms.RandomResizedCrop.get_params`` followed by ``RandomHorizontalFlip``'s coin, call for call on
torch's generator, so that with the same seed it yields the very boxes and flips the reference pipeline would
(``tests/test_augment_cpu.py`` checks that against torchvision itself and against the golden vectors).
``random_resized_crop_params_batched`` draws a whole batch with a handful of vectorised calls (same distribution,
different stream) for when 256 x 5 tiny generator calls per step would show on the host thread.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .shards import rank_indices

SCALE = (0.5, 1.0)                    # main_direct.py:160,166
RATIO = (3.0 / 4.0, 4.0 / 3.0)        # torchvision default


def _central_crop(height: int, width: int, ratio) -> Tuple[int, int, int, int]:
    """The fallback of get_params after ten rejected attempts: the whole image, clamped to the ratio bounds."""
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def random_resized_crop_params(n: int, height: int, width: int, scale: Sequence[float] = SCALE,
                               ratio: Sequence[float] = RATIO, flip_p: float = 0.5,
                               generator: Optional[torch.Generator] = None):
    """``(boxes int32 [n,4] = (top, left, h, w), flips uint8 [n])`` for ``n`` consecutive samples.

    Draw for draw what one ``__getitem__`` of the reference consumes: per sample up to ten
    ``(uniform area, uniform log-ratio)`` attempts, two ``randint`` for the position of the first box that fits,
    then ``torch.rand(1) < flip_p``.  ``generator=None`` uses the global generator, as torchvision does.
    """
    kw = {} if generator is None else {"generator": generator}
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    boxes = np.empty((n, 4), dtype=np.int32)
    flips = np.empty((n,), dtype=np.uint8)
    for s in range(n):
        box = None
        for _ in range(10):
            target_area = area * torch.empty(1).uniform_(scale[0], scale[1], **kw).item()
            aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], **kw)).item()
            w = int(round(math.sqrt(target_area * aspect_ratio)))
            h = int(round(math.sqrt(target_area / aspect_ratio)))
            if 0 < w <= width and 0 < h <= height:
                i = torch.randint(0, height - h + 1, size=(1,), **kw).item()
                j = torch.randint(0, width - w + 1, size=(1,), **kw).item()
                box = (i, j, h, w)
                break
        boxes[s] = box if box is not None else _central_crop(height, width, ratio)
        flips[s] = 1 if bool(torch.rand(1, **kw) < flip_p) else 0
    return boxes, flips


def random_resized_crop_params_batched(n: int, height: int, width: int, scale: Sequence[float] = SCALE,
                                       ratio: Sequence[float] = RATIO, flip_p: float = 0.5,
                                       generator: Optional[torch.Generator] = None):
    """Same distribution as ``random_resized_crop_params`` from five vectorised draws (ten candidate boxes per sample
    at once, the first that fits wins, the central crop otherwise); the stream of random numbers differs."""
    kw = {} if generator is None else {"generator": generator}
    area = float(height * width)
    lo, hi = math.log(ratio[0]), math.log(ratio[1])
    target = area * torch.empty(n, 10, dtype=torch.float64).uniform_(scale[0], scale[1], **kw)
    aspect = torch.exp(torch.empty(n, 10, dtype=torch.float64).uniform_(lo, hi, **kw))
    w = torch.round(torch.sqrt(target * aspect)).long()           # round-half-even vs Python's round: same rule
    h = torch.round(torch.sqrt(target / aspect)).long()
    ok = (w > 0) & (w <= width) & (h > 0) & (h <= height)
    first = torch.where(ok.any(1), ok.float().argmax(1), torch.zeros(n, dtype=torch.long))
    rows = torch.arange(n)
    w, h, any_ok = w[rows, first], h[rows, first], ok.any(1)
    ci, cj, ch, cw = _central_crop(height, width, ratio)
    h = torch.where(any_ok, h, torch.full_like(h, ch))
    w = torch.where(any_ok, w, torch.full_like(w, cw))
    u = torch.rand(n, 2, dtype=torch.float64, **kw)
    i = torch.minimum((u[:, 0] * (height - h + 1)).long(), height - h)
    j = torch.minimum((u[:, 1] * (width - w + 1)).long(), width - w)
    i = torch.where(any_ok, i, torch.full_like(i, ci))
    j = torch.where(any_ok, j, torch.full_like(j, cj))
    boxes = torch.stack([i, j, h, w], 1).to(torch.int32).numpy()
    flips = (torch.rand(n, **kw) < flip_p).to(torch.uint8).numpy()
    return boxes, flips


def check_boxes(boxes: np.ndarray, height: int, width: int, size: Tuple[int, int], allow_downscale: bool = False):
    """Host-side validation before the draws are uploaded: boxes inside the image and -- unless the caller accepts the
    plain (non-antialiased) bilinear filter -- not larger than the output (see include/oodfq_b200.h)."""
    b = np.asarray(boxes)
    if b.ndim != 2 or b.shape[1] != 4:
        raise ValueError(f"boxes must be [N,4] (top, left, h, w), got {b.shape}")
    top, left, h, w = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    if (h < 1).any() or (w < 1).any() or (top < 0).any() or (left < 0).any() or (top + h > height).any() or \
            (left + w > width).any():
        raise ValueError("a crop box leaves the image")
    if not allow_downscale and ((h > size[0]).any() or (w > size[1]).any()):
        raise ValueError("a crop box is larger than the output: torchvision would antialias this down-scaling; pass "
                         "allow_downscale=True to accept the plain bilinear filter")


class _CropResizeFlip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, index, boxes, flips, size, channels, channels_last):
        ctx.save_for_backward(index, boxes, flips)
        ctx.like = images.detach()                   # shape / layout only; the backward never reads its values
        return ops.crop_resize_flip(images.detach(), index, boxes, flips, size, channels=channels,
                                    channels_last=channels_last)

    @staticmethod
    def backward(ctx, grad_out):
        index, boxes, flips = ctx.saved_tensors
        return ops.crop_resize_flip_backward(grad_out, ctx.like, index, boxes, flips), None, None, None, None, None, None


def crop_resize_flip(images, index, boxes, flips, size=None, channels=None, channels_last=True):
    """Differentiable ``ops.crop_resize_flip``: the gradient reaches ``images`` through one scatter kernel, as
    autograd reaches ``gaussian_data`` through torchvision's ``RHF(RRC(gaussian_data[j]))`` in the distillation loop
    (data_generate/distill_data.py:197-227).  ``size`` defaults to the image size."""
    if size is None:
        size = tuple(images.shape[2:])
    return _CropResizeFlip.apply(images, index, boxes, flips, size, channels, channels_last)


def batch_augmenter(channels_last: bool = True):
    """``(x, boxes, flips) -> augmented x`` for ``step.DistillStep(augment=...)``: sample ``j`` is the crop
    ``boxes[j]`` of ``x[j]`` resized back to the image size and mirrored when ``flips[j]``, differentiable, one kernel
    each way for the whole batch (the reference loops over the images with torchvision,
    data_generate/distill_data.py:205-210)."""
    def run(x, boxes, flips):
        dev = x.device
        index = torch.arange(x.shape[0], device=dev)
        b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32)).to(dev, non_blocking=True)
        f = torch.from_numpy(np.ascontiguousarray(flips, dtype=np.uint8)).to(dev, non_blocking=True)
        check_boxes(boxes, x.shape[2], x.shape[3], tuple(x.shape[2:]))
        return crop_resize_flip(x, index, b, f, channels=x.shape[1], channels_last=channels_last)
    return run


class DeviceShards:
    """The shard set resident on one GPU, iterated as augmented per-rank batches ``(images, labels)`` on the device.

    ``images`` ``[M,C,H,W]`` float32 / ``labels`` ``[M]`` int64 as ``shards.load_shards`` returns them; the split over
    ranks and epochs is ``shards.rank_indices`` (DistributedSampler semantics).  ``augment=False`` yields the plain
    gather (identity boxes, no flips) through the same kernel.  Batches are written into ``slots`` rotating output
    buffers, so a consumer may hold ``slots - 1`` earlier batches; ``drop_last`` defaults to True because the QAT
    step replayed as a CUDA graph needs a fixed batch shape.
    """

    def __init__(self, images: np.ndarray, labels: np.ndarray, batch: int, device, rank: int = 0, world: int = 1,
                 shuffle: bool = True, seed: int = 0, size=None, scale: Sequence[float] = SCALE,
                 ratio: Sequence[float] = RATIO, flip_p: float = 0.5, augment: bool = True,
                 channels_last: bool = True, drop_last: bool = True, slots: int = 2, exact_stream: bool = False):
        if batch <= 0:
            raise ValueError("DeviceShards: batch must be positive")
        images = np.ascontiguousarray(images, dtype=np.float32)
        if images.ndim != 4 or images.shape[1] not in (1, 3):
            raise ValueError(f"DeviceShards: expected [M,1|3,H,W] images, got {images.shape}")
        if len(images) != len(labels):
            raise ValueError(f"DeviceShards: {len(images)} images but {len(labels)} labels")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ood_dfq_b200: DeviceShards keeps the image set in GPU memory; there is no CPU path "
                               "(use shards.ShardBatches for host-side batches)")
        self.images = torch.from_numpy(images).to(self.device)
        self.labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int64)).to(self.device)
        self.batch, self.rank, self.world = batch, rank, world
        self.shuffle, self.seed, self.drop_last, self.augment = shuffle, seed, drop_last, augment
        self.scale, self.ratio, self.flip_p, self.exact_stream = tuple(scale), tuple(ratio), flip_p, exact_stream
        h, w = images.shape[2:]
        self.size = (h, w) if size is None else ((size, size) if isinstance(size, int) else tuple(size))
        self.channels_last = channels_last
        self.epoch = 0
        self.generator = torch.Generator()
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self._out = [torch.empty((batch, 3, *self.size), dtype=torch.float32, device=self.device, memory_format=fmt)
                     for _ in range(max(1, slots))]
        # pinned staging for the per-step draws: index (8 B), box (16 B) and flip (1 B) per sample
        self._stage = [(torch.empty(batch, dtype=torch.int64).pin_memory(),
                        torch.empty((batch, 4), dtype=torch.int32).pin_memory(),
                        torch.empty(batch, dtype=torch.uint8).pin_memory()) for _ in range(max(1, slots))]
        self._copied = [None] * len(self._stage)        # event: the slot's staged draws have reached the device
        self._next = 0

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    @property
    def sampler(self):
        """``direct_dataload.sampler.set_epoch(epoch)`` (trainer_direct.py:447) keeps working on this object."""
        return self

    def __len__(self):
        per_rank = -(-len(self.labels) // self.world)
        return per_rank // self.batch if self.drop_last else -(-per_rank // self.batch)

    def draw(self, n: int):
        """Boxes and flips of the next ``n`` samples of this rank (host arrays)."""
        h, w = self.images.shape[2:]
        if not self.augment:
            boxes = np.tile(np.array([0, 0, h, w], dtype=np.int32), (n, 1))
            return boxes, np.zeros(n, dtype=np.uint8)
        fn = random_resized_crop_params if self.exact_stream else random_resized_crop_params_batched
        boxes, flips = fn(n, h, w, self.scale, self.ratio, self.flip_p, self.generator)
        check_boxes(boxes, h, w, self.size)
        return boxes, flips

    def __iter__(self):
        idx = rank_indices(len(self.labels), self.rank, self.world, self.epoch, self.shuffle, self.seed)
        # per-(rank, epoch) stream, so ranks do not mirror each other's crops (the reference seeds every rank alike,
        # SURVEY.md section 8(e) "quirk")
        self.generator.manual_seed((self.seed * 1000003 + self.epoch) * 1009 + self.rank)
        for start in range(0, len(idx), self.batch):
            chunk = idx[start:start + self.batch]
            n = len(chunk)
            if n < self.batch and self.drop_last:
                return
            boxes, flips = self.draw(n)
            slot = self._next
            self._next = (self._next + 1) % len(self._out)
            h_idx, h_box, h_flip = self._stage[slot]
            if self._copied[slot] is not None:
                # the host may run far ahead of the GPU (graph replay): never overwrite staged draws whose
                # asynchronous copy has not executed yet
                self._copied[slot].synchronize()
            h_idx[:n].copy_(torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.int64)))
            h_box[:n].copy_(torch.from_numpy(boxes))
            h_flip[:n].copy_(torch.from_numpy(flips))
            d_idx = h_idx[:n].to(self.device, non_blocking=True)
            d_box = h_box[:n].to(self.device, non_blocking=True)
            d_flip = h_flip[:n].to(self.device, non_blocking=True)
            self._copied[slot] = torch.cuda.Event()
            self._copied[slot].record()
            out = self._out[slot][:n]
            ops.crop_resize_flip(self.images, d_idx, d_box, d_flip, self.size, channels=3,
                                 channels_last=self.channels_last, out=out)
            yield out, self.labels.index_select(0, d_idx)
