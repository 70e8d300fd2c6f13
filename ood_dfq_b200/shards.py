"""Reader for the reference's synthetic-image shards: the data format on the input side of the QAT step.

``data_generate/generate_data.py:1064-1074`` writes, for each of four groups ``g = 1..4``,

    {data_prefix}{g}.pickle     pickle of  [np.ndarray[B, C, H, W] float32, ...]
    {label_prefix}{g}.pickle    pickle of  [np.ndarray[B] int64, ...]

and ``direct_dataset`` (``main_direct.py:150-207``) concatenates every array of every group, wraps the result in a
``Dataset`` and lets a ``DistributedSampler`` split it over the ranks (``main_direct.py:525-527``).  This module
restates that host-side logic without torchvision / the logger: ``load_shards`` is the concatenation,
``ShardBatches`` the per-rank batch stream -- whole batches gathered into reusable (pinned, when a GPU is present)
host buffers in the memory format the kernels run in, which is what the double-buffered H2D prefetcher of
``bench.py`` consumes.  The per-sample ``RandomResizedCrop`` / flip of the reference (``main_direct.py:158-169``) is a
caller-supplied ``transform`` here; the device-side version (image set resident in HBM, one kernel per batch) is
``ood_dfq_b200/augment.py::DeviceShards``.

The files are pickles: like the reference, only load shards you produced yourself.
"""
from __future__ import annotations

import pickle
from typing import Callable, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch


def load_shards(data_prefix: str, label_prefix: str, groups: Iterable[int] = (1, 2, 3, 4)) -> Tuple[np.ndarray, np.ndarray]:
    """``(images [M,C,H,W] float32, labels [M] int64)``: every array of every group, concatenated in file order
    (``main_direct.py:173-195``)."""
    data, labels = [], []
    for g in groups:
        with open(f"{data_prefix}{g}.pickle", "rb") as fp:
            data.extend(np.asarray(a) for a in pickle.load(fp))
        with open(f"{label_prefix}{g}.pickle", "rb") as fp:
            labels.extend(np.asarray(a) for a in pickle.load(fp))
    if not data:
        raise ValueError("load_shards: no groups given")
    images = np.concatenate(data, axis=0)
    targets = np.concatenate(labels, axis=0)
    if len(images) != len(targets):                      # main_direct.py:197
        raise ValueError(f"load_shards: {len(images)} images but {len(targets)} labels")
    if images.ndim != 4:
        raise ValueError(f"load_shards: expected [M,C,H,W] images, got shape {images.shape}")
    return np.ascontiguousarray(images, dtype=np.float32), np.ascontiguousarray(targets, dtype=np.int64)


def write_shards(data_prefix: str, label_prefix: str, images: np.ndarray, labels: np.ndarray, groups: int = 4):
    """Write ``images`` / ``labels`` in the reference's layout (``generate_data.py:1064-1074``), split evenly over
    ``groups`` files -- for tests and for feeding the reference itself."""
    parts = np.array_split(np.arange(len(images)), groups)
    for g, idx in enumerate(parts, start=1):
        with open(f"{data_prefix}{g}.pickle", "wb") as fp:
            pickle.dump([np.ascontiguousarray(images[idx], dtype=np.float32)], fp, protocol=pickle.HIGHEST_PROTOCOL)
        with open(f"{label_prefix}{g}.pickle", "wb") as fp:
            pickle.dump([np.ascontiguousarray(labels[idx], dtype=np.int64)], fp, protocol=pickle.HIGHEST_PROTOCOL)


def rank_indices(n: int, rank: int, world: int, epoch: int = 0, shuffle: bool = True, seed: int = 0) -> np.ndarray:
    """The sample indices ``DistributedSampler(dataset, world, rank)`` hands rank ``rank`` in ``epoch``: a seeded
    permutation (``seed + epoch``), padded by wrapping to a multiple of ``world``, strided by ``world``."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n, generator=g).numpy()
    else:
        order = np.arange(n)
    total = -(-n // world) * world
    if total > n:
        reps = -(-(total - n) // n)
        order = np.concatenate([order] + [order] * reps)[:total]
    return order[rank:total:world]


class ShardBatches:
    """Per-rank batches of a shard set as ``(images [B,3,H,W], labels [B])`` host tensors.

    One-channel images are repeated to three channels (the ``Lambda`` of ``main_direct.py:161,167``); ``transform``
    (a callable on one ``[C,H,W]`` tensor, e.g. the reference's torchvision pipeline) runs per sample when given.
    Batches are gathered into ``slots`` reusable buffers -- pinned when CUDA is available, channels_last by default --
    so the consumer can issue ``non_blocking`` H2D copies from them; a buffer is reused after ``slots`` further
    batches.  ``drop_last`` defaults to True: the QAT step captured as a CUDA graph needs a fixed batch shape.
    """

    def __init__(self, images: np.ndarray, labels: np.ndarray, batch: int, rank: int = 0, world: int = 1,
                 shuffle: bool = True, seed: int = 0, transform: Optional[Callable] = None, drop_last: bool = True,
                 channels_last: bool = True, slots: int = 3, out_size: Optional[Sequence[int]] = None):
        if batch <= 0:
            raise ValueError("ShardBatches: batch must be positive")
        self.images, self.labels = images, labels
        self.batch, self.rank, self.world = batch, rank, world
        self.shuffle, self.seed, self.transform, self.drop_last = shuffle, seed, transform, drop_last
        self.epoch = 0
        c, h, w = images.shape[1:]
        if out_size is not None:
            h, w = out_size
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        pin = torch.cuda.is_available()
        self._slots = []
        for _ in range(max(1, slots)):
            buf = torch.empty((batch, 3 if c == 1 else c, h, w), dtype=torch.float32).contiguous(memory_format=fmt)
            lab = torch.empty((batch,), dtype=torch.int64)
            self._slots.append((buf.pin_memory() if pin else buf, lab.pin_memory() if pin else lab))
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

    def __iter__(self):
        idx = rank_indices(len(self.labels), self.rank, self.world, self.epoch, self.shuffle, self.seed)
        for start in range(0, len(idx), self.batch):
            chunk = idx[start:start + self.batch]
            if len(chunk) < self.batch and self.drop_last:
                return
            buf, lab = self._slots[self._next]
            self._next = (self._next + 1) % len(self._slots)
            n = len(chunk)
            if self.transform is None:
                src = torch.from_numpy(self.images[chunk])
                buf[:n].copy_(src.expand(-1, buf.shape[1], -1, -1) if src.shape[1] == 1 else src)
            else:
                for j, i in enumerate(chunk):
                    x = torch.from_numpy(self.images[i])
                    x = self.transform(x)
                    buf[j].copy_(x.repeat(3, 1, 1) if x.size(0) == 1 else x)
            lab[:n].copy_(torch.from_numpy(self.labels[chunk]))
            yield buf[:n], lab[:n]
