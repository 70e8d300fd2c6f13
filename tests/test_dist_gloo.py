"""Data-parallel host logic on CPU: world_size-2 gloo process groups (no GPU needed).

What is covered: the packed activation-range all-reduce (reduce_minmax), the flat gradient
all-reduce of the QAT step against the single-process global-batch step, and the additivity of the
shifted BN partial sums that BNStatLoss(sync=True) all-reduces.  The modules are the CPU oracle's
(the CUDA mirror cannot run here); the collective code paths are the product's.
"""
import copy
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bns_torch, fq_torch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):          # drain BEFORE join: a child blocks in put() until its payload is read
        r, v = q.get()
        if isinstance(v, BaseException):
            for p in procs:
                p.kill()
            raise v
        out[r] = v
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    return out


def _entry(fn, rank, world, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    except BaseException as e:      # surface the failure in the parent instead of hanging its q.get()
        q.put((rank, RuntimeError(f"rank {rank}: {type(e).__name__}: {e}")))
        raise
    finally:
        dist.destroy_process_group()


def _tiny_student(seed=1):
    from ood_dfq_b200 import nets, surgery
    torch.manual_seed(seed)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=fq_torch)
    return teacher, student


# ------------------------------------------------------------------ reduce_minmax
def _w_reduce_minmax(rank, world):
    from ood_dfq_b200 import dist as ddist
    _, student = _tiny_student()
    g = torch.Generator().manual_seed(100 + rank)
    with torch.no_grad():
        for _ in range(2):
            student(torch.randn(4, 3, 32, 32, generator=g))
    acts = [m for m in student.modules() if isinstance(m, fq_torch.OracleQuantAct)]
    before = torch.stack([torch.cat([m.x_min, m.x_max]) for m in acts])
    ddist.reduce_minmax(student, act_types=(fq_torch.OracleQuantAct,))
    after = torch.stack([torch.cat([m.x_min, m.x_max]) for m in acts])
    assert all(m.x_min.shape == (1,) and "x_min" in dict(m.named_buffers()) for m in acts)
    return before.numpy(), after.numpy()


def test_reduce_minmax_is_one_packed_mean():
    out = _spawn(_w_reduce_minmax, 2)
    b0, a0 = out[0]
    b1, a1 = out[1]
    assert b0.shape[0] == 19 and not np.array_equal(b0, b1)          # 19 QuantAct sites, rank-local ranges
    np.testing.assert_array_equal(a0, a1)                             # identical on every rank afterwards
    np.testing.assert_array_equal(a0, ((b0 + b1) / np.float32(2)).astype(np.float32))   # (sum)/world, fp32


# ------------------------------------------------------------------ QAT step: 2 ranks == global batch
def _make_step(student, teacher):
    from ood_dfq_b200 import nets, step
    return step.QATStep(student, teacher, lr=1e-2, momentum=0.9, weight_decay=1e-4, temperature=20.0, alpha=20.0,
                        lam=1000.0, eps=0.01, unit_types=(nets.ResUnit,))


def _calibrated_pair():
    teacher, student = _tiny_student()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        student(torch.randn(8, 3, 32, 32, generator=g))
    for m in student.modules():
        if isinstance(m, fq_torch.OracleQuantAct):
            m.fix()
    return teacher, student


def _global_batch():
    return torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(42))


def _w_qat_step(rank, world):
    from ood_dfq_b200 import dist as ddist
    teacher, student = _calibrated_pair()
    qat = _make_step(student, teacher)
    qat(ddist.shard_batch(_global_batch(), rank, world))
    return {n: p.detach().numpy().copy() for n, p in student.named_parameters()}


def test_two_rank_step_matches_global_batch_step():
    out = _spawn(_w_qat_step, 2)
    teacher, student = _calibrated_pair()
    before = {n: p.detach().clone() for n, p in student.named_parameters()}
    _make_step(student, teacher)(_global_batch())
    moved = 0
    for n, p in student.named_parameters():
        np.testing.assert_array_equal(out[0][n], out[1][n])            # replicas stay in lock-step
        ref = p.detach().numpy()
        scale = max(1e-12, float(np.abs(ref - before[n].numpy()).max()))
        # the update of the 2-rank run equals the global-batch update (the sign perturbation is per-image,
        # KD is a batch mean, FA a mean over batch x channels: all average correctly over equal shards)
        assert np.abs(out[0][n] - ref).max() <= 2e-3 * scale + 1e-7, n
        moved += float(np.abs(ref - before[n].numpy()).max() > 0)
    assert moved > 10


# ------------------------------------------------------------------ BN partial sums are additive
def _w_bn_sums(rank, world):
    from ood_dfq_b200 import dist as ddist
    x = ddist.shard_batch(torch.randn(12, 6, 7, 7, generator=torch.Generator().manual_seed(3)) * 2 + 5, rank, world)
    shift = torch.linspace(4.5, 5.5, 6)                                # the (replicated) BN running mean
    d = x - shift.view(1, -1, 1, 1)
    sums = torch.cat([d.sum([0, 2, 3]), (d * d).sum([0, 2, 3])])       # what oodfq_bn_stats_forward emits
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)                        # what BNStatLoss(sync=True) does
    count = world * x.shape[0] * 49
    m1, m2 = sums[:6].double() / count, sums[6:].double() / count
    return (shift.double() + m1).float().numpy(), (m2 - m1 * m1).float().numpy()


def test_bn_partial_sums_allreduce_gives_global_statistics():
    out = _spawn(_w_bn_sums, 2)
    x = torch.randn(12, 6, 7, 7, generator=torch.Generator().manual_seed(3)) * 2 + 5
    mean, var = bns_torch.channel_stats(x)
    for r in (0, 1):
        np.testing.assert_allclose(out[r][0], mean.numpy(), rtol=1e-6)
        np.testing.assert_allclose(out[r][1], var.numpy(), rtol=1e-5)


# ------------------------------------------------------------------ BNStatLoss(sync=True): the product's own manager
def _sync_net():
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 12, 3, stride=2, padding=1, bias=False), torch.nn.BatchNorm2d(12)).eval()
    g = torch.Generator().manual_seed(5)
    for m in net:
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return net


def _sync_batch():
    return torch.randn(8, 3, 10, 10, generator=torch.Generator().manual_seed(6)) * 1.5 + 0.3


def _w_bns_sync(rank, world):
    """``bns.BNStatLoss(sync=True)`` itself -- hooks, packed fp64 sums, the all-reduce, the fused backward -- with its
    kernel launches swapped for oracle arithmetic (tests/cpu_ops_shim.py, test-only)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_shim
    from ood_dfq_b200 import bns, dist as ddist
    net = _sync_net()
    x = ddist.shard_batch(_sync_batch(), rank, world).clone().requires_grad_(True)
    with cpu_ops_shim.installed():
        mgr = bns.BNStatLoss(net, sync=True)
        net(x)
        loss = mgr.loss()
        loss.backward()
    return loss.item(), x.grad.numpy().copy()


def test_synced_bn_stat_loss_is_the_global_batch_loss():
    out = _spawn(_w_bns_sync, 2)
    net = _sync_net()
    x = _sync_batch().requires_grad_(True)
    tap = bns_torch.StatTap(net)
    net(x)
    loss = tap.loss("trainer")
    loss.backward()
    assert abs(out[0][0] - loss.item()) <= 1e-6 * abs(loss.item()) and out[0][0] == out[1][0]
    # every rank holds d(global loss)/d(its own images): together the global gradient
    got = np.concatenate([out[0][1], out[1][1]])
    np.testing.assert_allclose(got, x.grad.numpy(), rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------ packed reduce_minmax == the reference's own code
_REF_TRAINER = "/root/reference/trainer_direct.py"


def _w_reduce_minmax_vs_reference(rank, world):
    """Each rank calibrates on its own data; then the reference's ``Trainer.reduce_minmax`` (its source, lines 368-374,
    2 all-reduces per QuantAct) and the packed one-call version are applied to copies of the same MIRROR student."""
    import textwrap
    import types
    import ood_dfq_b200
    from ood_dfq_b200 import dist as ddist, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    ood_dfq_b200.install()
    ns = {"dist": dist}
    exec("from quantization_utils.quant_modules import *", ns)
    with open(_REF_TRAINER) as f:
        exec(compile(textwrap.dedent("".join(f.readlines()[367:374]).expandtabs(4)), _REF_TRAINER, "exec"), ns)
    torch.manual_seed(1)
    student = surgery.quantize_model(nets.resnet20_cifar(num_classes=10), 4, 4)      # mirror classes, CPU buffers only
    g = torch.Generator().manual_seed(200 + rank)
    for m in student.modules():
        if isinstance(m, qm.QuantAct):
            m.x_min.copy_(-torch.rand(1, generator=g))
            m.x_max.copy_(torch.rand(1, generator=g) * 5)
    theirs, ours = copy.deepcopy(student), copy.deepcopy(student)
    ns["reduce_minmax"](types.SimpleNamespace(model=types.SimpleNamespace(module=theirs)))
    ddist.reduce_minmax(ours)
    pack = lambda net: torch.stack([torch.cat([m.x_min, m.x_max]) for m in net.modules() if isinstance(m, qm.QuantAct)])
    return pack(theirs).numpy(), pack(ours).numpy(), pack(student).numpy()


@pytest.mark.skipif(not os.path.isfile(_REF_TRAINER), reason="reference tree not mounted (GPU box)")
def test_packed_reduce_minmax_equals_the_reference_source_on_two_ranks():
    out = _spawn(_w_reduce_minmax_vs_reference, 2)
    for r in (0, 1):
        theirs, ours, _ = out[r]
        np.testing.assert_array_equal(theirs, ours)                     # bit-identical to the reference's own code
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][1], ((out[0][2] + out[1][2]) / np.float32(2)).astype(np.float32))
    assert not np.array_equal(out[0][2], out[1][2])
