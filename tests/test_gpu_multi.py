"""Multi-GPU parity (NCCL, one process per GPU): needs >= 2 CUDA devices, skipped otherwise.

Oracle for the N-rank result = the same computation in ONE process on the concatenated global batch
(SURVEY.md section 8(e)).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(fn, rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        q.put((rank, fn(rank, world)))
    except BaseException as e:
        q.put((rank, RuntimeError(f"rank {rank}: {type(e).__name__}: {e}")))
        raise
    finally:
        dist.destroy_process_group()


def _spawn(fn, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        r, v = q.get()
        if isinstance(v, BaseException):
            for p in procs:
                p.kill()
            raise v
        out[r] = v
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return out


def _net():
    torch.manual_seed(5)
    nn = torch.nn
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1, bias=False), nn.BatchNorm2d(8), nn.ReLU(),
                        nn.Conv2d(8, 16, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(16)).eval()
    g = torch.Generator().manual_seed(6)
    for m in net:
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return net


def _batch():
    return torch.randn(8, 3, 14, 14, generator=torch.Generator().manual_seed(9))


def _w_bns(rank, world):
    from ood_dfq_b200 import bns, dist as ddist
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", rank)
    net = _net().to(dev)
    stat = bns.BNStatLoss(net, sync=True)
    x = ddist.shard_batch(_batch(), rank, world).to(dev).requires_grad_(True)
    net(x)
    loss = stat.loss()
    loss.backward()
    return loss.item(), x.grad.cpu().numpy()


def test_bns_loss_two_ranks_equals_global_batch():
    out = _spawn(_w_bns, 2)
    from oracle import bns_torch
    net = _net()
    tap = bns_torch.StatTap(net)
    x = _batch().requires_grad_(True)
    net(x)
    loss = tap.loss("trainer")
    loss.backward()
    for r in (0, 1):
        np.testing.assert_allclose(out[r][0], loss.item(), rtol=1e-5)
        np.testing.assert_allclose(out[r][1], x.grad[4 * r: 4 * r + 4].numpy(), rtol=1e-4, atol=1e-9)


def _w_minmax(rank, world):
    from ood_dfq_b200 import dist as ddist
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    dev = torch.device("cuda", rank)
    model = torch.nn.Sequential(qm.QuantAct(4), qm.QuantAct(4), qm.QuantAct(8)).to(dev)
    g = torch.Generator().manual_seed(20 + rank)
    for _ in range(2):
        model(torch.relu(torch.randn(4, 8, 6, 6, generator=g)).to(dev))
    before = torch.stack([torch.cat([m.x_min, m.x_max]) for m in model]).cpu().numpy()
    ddist.reduce_minmax(model)
    after = torch.stack([torch.cat([m.x_min, m.x_max]) for m in model]).cpu().numpy()
    y = model(torch.ones(1, 1, 2, 2, device=dev))          # still usable (frozen or not) after re-assignment
    return before, after, float(y.sum())


def test_reduce_minmax_nccl():
    out = _spawn(_w_minmax, 2)
    b0, a0, _ = out[0]
    b1, a1, _ = out[1]
    np.testing.assert_array_equal(a0, a1)
    np.testing.assert_array_equal(a0, ((b0 + b1) / np.float32(2)).astype(np.float32))
