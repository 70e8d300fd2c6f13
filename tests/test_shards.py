"""Reader of the reference's pickle shards (generate_data.py:1064-1074, main_direct.py:150-207)."""
import pickle

import numpy as np
import pytest
import torch

from ood_dfq_b200 import shards


def _make(tmp_path, n=37, c=3, hw=8, seed=0):
    rng = np.random.default_rng(seed)
    images = rng.standard_normal((n, c, hw, hw)).astype(np.float32)
    labels = rng.integers(0, 10, size=n).astype(np.int64)
    dp, lp = str(tmp_path / "data_"), str(tmp_path / "label_")
    shards.write_shards(dp, lp, images, labels)
    return dp, lp, images, labels


def test_layout_is_the_references(tmp_path):
    dp, lp, images, labels = _make(tmp_path)
    with open(dp + "1.pickle", "rb") as fp:
        first = pickle.load(fp)
    assert isinstance(first, list) and len(first) == 1 and first[0].dtype == np.float32 and first[0].ndim == 4
    # what main_direct.py:173-195 does with the files
    ref_data = np.concatenate([np.concatenate(pickle.load(open(f"{dp}{i}.pickle", "rb")), axis=0) for i in range(1, 5)])
    ref_lab = np.concatenate([np.concatenate(pickle.load(open(f"{lp}{i}.pickle", "rb")), axis=0) for i in range(1, 5)])
    got_data, got_lab = shards.load_shards(dp, lp)
    assert np.array_equal(got_data, ref_data) and np.array_equal(got_lab, ref_lab)
    assert np.array_equal(got_data, images) and np.array_equal(got_lab, labels)


def test_several_arrays_per_file_and_mismatch(tmp_path):
    a = np.zeros((2, 1, 4, 4), np.float32)
    b = np.ones((3, 1, 4, 4), np.float32)
    dp, lp = str(tmp_path / "d"), str(tmp_path / "l")
    for g in (1, 2):
        pickle.dump([a, b], open(f"{dp}{g}.pickle", "wb"))
        pickle.dump([np.zeros(2, np.int64), np.ones(3, np.int64)], open(f"{lp}{g}.pickle", "wb"))
    data, lab = shards.load_shards(dp, lp, groups=(1, 2))
    assert data.shape == (10, 1, 4, 4) and lab.tolist() == [0, 0, 1, 1, 1] * 2
    pickle.dump([np.zeros(4, np.int64)], open(f"{lp}2.pickle", "wb"))
    with pytest.raises(ValueError):
        shards.load_shards(dp, lp, groups=(1, 2))


def test_rank_split_matches_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    data = list(range(37))
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=5)
            ref.set_epoch(3)
            got = shards.rank_indices(37, rank, world, epoch=3, shuffle=True, seed=5)
            assert list(ref) == got.tolist()
            seen.extend(got.tolist())
        assert set(seen) == set(range(37))
        ref = DistributedSampler(data, num_replicas=world, rank=0, shuffle=False)
        assert list(ref) == shards.rank_indices(37, 0, world, shuffle=False).tolist()


def test_batches_cover_the_rank_share(tmp_path):
    dp, lp, images, labels = _make(tmp_path, n=40)
    data, lab = shards.load_shards(dp, lp)
    it = shards.ShardBatches(data, lab, batch=6, rank=1, world=2, seed=1, channels_last=True, drop_last=False)
    it.set_epoch(2)
    want = shards.rank_indices(40, 1, 2, epoch=2, seed=1)
    got_x, got_y = [], []
    for x, y in it:
        assert x.dtype == torch.float32 and x.shape[1:] == (3, 8, 8)
        got_x.append(x.clone())
        got_y.append(y.clone())
    assert len(got_x) == len(it) == 4
    assert torch.equal(torch.cat(got_x), torch.from_numpy(images[want])) and torch.equal(torch.cat(got_y), torch.from_numpy(labels[want]))
    full = shards.ShardBatches(data, lab, batch=6, rank=1, world=2, seed=1)
    shapes = [x.shape[0] for x, _ in full]
    assert shapes == [6, 6, 6] and len(full) == 3                      # drop_last: fixed shape for graph replay
    x0, _ = next(iter(full))
    assert x0.is_contiguous(memory_format=torch.channels_last)


def test_grayscale_and_transform(tmp_path):
    dp, lp, images, labels = _make(tmp_path, n=8, c=1)
    data, lab = shards.load_shards(dp, lp)
    plain = shards.ShardBatches(data, lab, batch=4, shuffle=False)
    x, y = next(iter(plain))
    assert x.shape == (4, 3, 8, 8) and torch.equal(x[:, 0], x[:, 2]) and torch.equal(x[:, 0], torch.from_numpy(images[:4, 0]))
    flipped = shards.ShardBatches(data, lab, batch=4, shuffle=False, transform=lambda t: t.flip(-1))
    xf, _ = next(iter(flipped))
    assert torch.equal(xf, x.flip(-1))
    crop = shards.ShardBatches(data, lab, batch=4, shuffle=False, transform=lambda t: t[:, :4, :4], out_size=(4, 4))
    xc, _ = next(iter(crop))
    assert xc.shape == (4, 3, 4, 4) and torch.equal(xc, x[:, :, :4, :4])


def test_batch_stream_stands_in_for_the_dataloader_interface(tmp_path):
    """What ``Trainer.train`` does with ``direct_dataload`` (trainer_direct.py:446-448, :492-497): ``.sampler.set_epoch``,
    ``iter`` / ``next`` with re-iteration on exhaustion, ``len``."""
    from ood_dfq_b200 import shards
    rng = np.random.default_rng(1)
    images, labels = rng.standard_normal((10, 3, 8, 8)).astype(np.float32), np.arange(10, dtype=np.int64)
    stream = shards.ShardBatches(images, labels, batch=4, seed=3)
    stream.sampler.set_epoch(2)
    assert stream.epoch == 2 and len(stream) == 2
    first = [y.clone() for _, y in stream]
    stream.sampler.set_epoch(3)
    second = [y.clone() for _, y in stream]
    assert len(first) == len(second) == 2 and not all(torch.equal(a, b) for a, b in zip(first, second))
    it = iter(stream)
    next(it), next(it)
    with pytest.raises(StopIteration):
        next(it)
