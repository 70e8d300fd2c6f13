"""CPU restatement of the reference's per-sample augmentation pipeline.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference builds, in ``direct_dataset.__init__``
(``/root/reference/main_direct.py:158-169``),

    Compose([RandomResizedCrop(size=img_size, scale=(0.5, 1.0)),
             Lambda(lambda x: x.repeat(3, 1, 1) if x.size(0) == 1 else x),
             RandomHorizontalFlip()])

and applies it to one ``[C,H,W]`` float tensor per ``__getitem__`` (``main_direct.py:200-204``).  The arithmetic
lives in a third-party dependency that is not under ``/root/reference``: **torchvision** (un-pinned by the
reference -- its README only pins ``Pytorch == 1.8.1``; this image has torchvision 0.26.0 on torch 2.11).  Its
published algorithm for tensors, restated below with plain torch calls:

* ``RandomResizedCrop.get_params`` (torchvision/transforms/transforms.py): up to ten attempts of
  ``area * U(scale)``, ``exp(U(log ratio))``, ``w = round(sqrt(area * ratio))``, ``h = round(sqrt(area / ratio))``,
  accepted when the box fits, then ``i = randint(0, H-h+1)``, ``j = randint(0, W-w+1)``; otherwise a central crop
  clamped to the ratio bounds.  Draws come from torch's global generator, one ``torch.empty(1).uniform_`` /
  ``torch.randint(size=(1,))`` per value, in that order.
* ``F.resized_crop`` = ``img[..., i:i+h, j:j+w]`` followed by ``torch.nn.functional.interpolate(mode="bilinear",
  align_corners=False, antialias=True)``.  The crop is never larger than the output here (``size`` is the image
  size), and for an up-scaling the antialiased filter has support 1, i.e. it IS the plain bilinear filter; the two
  ATen kernels differ in the last bit only (checked when the golden vectors are generated).
* ``RandomHorizontalFlip.forward``: ``torch.rand(1) < p`` then ``img.flip(-1)``.

Pinned by ``tests/golden/augment.npz``, which ``tools/make_golden.py`` writes by running torchvision's own
transform objects (not this file) at fixed seeds.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def get_params(height: int, width: int, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """``RandomResizedCrop.get_params``: ``(top, left, h, w)``, consuming the global generator like torchvision."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
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


def draw(height: int, width: int, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), p: float = 0.5):
    """The draws of one ``__getitem__``: the crop box, then the flip decision."""
    box = get_params(height, width, scale, ratio)
    return box, bool(torch.rand(1) < p)


def resized_crop_flip(img: torch.Tensor, box, flip: bool, size, antialias: bool = True) -> torch.Tensor:
    """One sample ``[C,H,W]`` through crop -> bilinear resize -> repeat -> flip (computed in ``img.dtype``)."""
    i, j, h, w = (int(v) for v in box)
    crop = img[..., i:i + h, j:j + w]
    out = F.interpolate(crop[None], size=tuple(size), mode="bilinear", align_corners=False, antialias=antialias)[0]
    if out.size(0) == 1:
        out = out.repeat(3, 1, 1)
    return out.flip(-1) if flip else out


def batch(images: torch.Tensor, index, boxes, flips, size, channels: int = 3, dtype=None) -> torch.Tensor:
    """A batch ``[N, channels, *size]`` from the image set ``[M,C,H,W]``; ``channels=1`` keeps grey images grey.
    Differentiable w.r.t. ``images`` (autograd runs ATen's interpolate backward, as it does for the reference's
    ``RHF(RRC(gaussian_data[j]))``, data_generate/distill_data.py:197-227)."""
    outs = []
    for n, m in enumerate(index):
        img = images[int(m)]
        img = img.to(dtype) if dtype is not None else img
        y = resized_crop_flip(img, boxes[n], bool(flips[n]), size)
        outs.append(y[:1] if channels == 1 else y)
    return torch.stack(outs)
