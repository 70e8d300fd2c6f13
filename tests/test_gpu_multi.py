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


def _qat_pair(dev):
    """A small W4A4 student / teacher with frozen ranges, identical on every caller (same seeds, same calibration)."""
    import copy
    from ood_dfq_b200 import nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.manual_seed(11)
    teacher = nets.perturb_bn_stats(nets.resnet20_cifar(num_classes=10))
    student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=qm)
    teacher, student = teacher.to(dev), student.to(dev)
    calib = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(12)).to(dev)
    with torch.no_grad():
        for _ in range(3):
            student(calib)
    surgery.freeze_model(student, qm)
    return teacher, student


def _qat_step(teacher, student):
    from ood_dfq_b200 import nets, step
    return step.QATStep(student, teacher, lr=1e-3, momentum=0.9, weight_decay=1e-4, temperature=20.0, alpha=20.0,
                        lam=1000.0, eps=0.01, unit_types=(nets.ResUnit,))


def _global_batches(n, steps):
    g = torch.Generator().manual_seed(13)
    return [torch.randn(n, 3, 32, 32, generator=g) for _ in range(steps)]


def _deterministic():
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _w_grad_exchange(rank, world):
    """Two SGD steps, each rank on its shard, gradients exchanged by ``FlatGrads.all_reduce_mean`` over NCCL."""
    from ood_dfq_b200 import dist as ddist
    _deterministic()
    dev = torch.device("cuda", rank)
    qat = _qat_step(*_qat_pair(dev))
    grads = []
    for glob in _global_batches(8 * world, 2):
        qat.compute(ddist.shard_batch(glob, rank, world).to(dev))
        qat.grads.all_reduce_mean(qat.group)
        grads.append(qat.grads.flat.cpu().numpy().copy())
        qat.opt.step()
    weights = torch.cat([p.detach().reshape(-1) for p in qat.student.parameters()]).cpu().numpy()
    return grads, weights


def test_gradient_exchange_nccl_equals_the_global_batch():
    """Row a15 (DDP of main_direct.py:484, backward_S trainer_direct.py:350-356) over NCCL: the all-reduced mean
    gradient of two ranks equals the gradient of the global batch walked shard by shard in ONE process with the
    exchange switched off, step after step; the updated weights are bit-identical on both ranks."""
    world = 2
    out = _spawn(_w_grad_exchange, world)
    from ood_dfq_b200 import dist as ddist
    _deterministic()
    dev = torch.device("cuda", 0)
    solo = _qat_step(*_qat_pair(dev))
    solo.exchange = False
    for it, glob in enumerate(_global_batches(8 * world, 2)):
        acc = torch.zeros_like(solo.grads.flat)
        for r in range(world):
            solo.compute(ddist.shard_batch(glob, r, world).to(dev))
            acc += solo.grads.flat
        want = (acc / world).cpu().numpy()
        for r in range(world):
            np.testing.assert_allclose(out[r][0][it], want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
        solo.grads.flat.copy_(acc / world)
        solo.opt.step()
    assert np.array_equal(out[0][1].view(np.int32), out[1][1].view(np.int32))          # replicas stay bit-identical
    want_w = torch.cat([p.detach().reshape(-1) for p in solo.student.parameters()]).cpu().numpy()
    np.testing.assert_allclose(out[0][1], want_w, rtol=1e-5, atol=1e-7)
