"""Host logic of the QAT step on CPU (oracle modules): the pruned final backward leaves the student's gradients
and update bit-identical to the reference's full ``loss.backward()`` sweep (trainer_direct.py:350-356)."""
import copy

import torch

from ood_dfq_b200 import nets, step, surgery
from oracle import fq_torch


def _pair():
    torch.manual_seed(1)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=fq_torch)
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn(4, 3, 32, 32, generator=g) for _ in range(2)]
    surgery.unfreeze_model(student, fq_torch)
    with torch.no_grad():
        student(xs[0])
    surgery.freeze_model(student, fq_torch)
    return teacher, student, xs


def test_pruned_backward_gives_the_same_update():
    results = []
    for prune in (True, False):
        teacher, student, xs = _pair()
        qat = step.QATStep(student, teacher, lr=1e-3, unit_types=(nets.ResUnit,), prune_backward=prune)
        losses = [qat(x).item() for x in xs]
        results.append((losses, [p.detach().clone() for p in student.parameters()], qat.grads.flat.clone()))
    (l0, p0, g0), (l1, p1, g1) = results
    assert l0 == l1
    assert torch.equal(g0, g1)
    assert all(torch.equal(a, b) for a, b in zip(p0, p1))


def test_teacher_stays_frozen_and_taps_line_up():
    teacher, student, xs = _pair()
    before = [p.detach().clone() for p in teacher.parameters()]
    qat = step.QATStep(student, teacher, lr=1e-3, unit_types=(nets.ResUnit,))
    qat(xs[0])
    assert all(torch.equal(a, b) for a, b in zip(before, teacher.parameters()))
    assert all(p.grad is None for p in teacher.parameters())
    assert len(qat.tap_s.maps) == len(qat.tap_t.maps) == 9


def test_device_side_plateau_scheduler_follows_torchs():
    """``PlateauOnDevice`` (the sync-free twin of the distillation loop's ReduceLROnPlateau, distill_data.py:185-188,
    :275) takes the same decisions as torch's scheduler on the same fp32 losses: improving, flat and worsening runs,
    the relative threshold, the floor at min_lr."""
    import random

    from ood_dfq_b200.step import PlateauOnDevice
    for seed in range(12):
        random.seed(seed)
        p = torch.zeros(1, requires_grad=True)
        opt = torch.optim.Adam([p], lr=0.5)
        sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, min_lr=1e-4, patience=4)
        lr = torch.tensor(0.5)
        dev = PlateauOnDevice(lr, patience=4)
        loss, reductions = 10.0, 0
        for it in range(300):
            r = random.random()
            loss *= 0.99 if r < 0.3 else (1.01 if r < 0.5 else (1 - 1e-4 if r < 0.55 else 1.0))
            l32 = torch.tensor(loss, dtype=torch.float32)
            before = opt.param_groups[0]["lr"]
            sch.step(l32.item())
            dev.step(l32)
            after = opt.param_groups[0]["lr"]
            reductions += int(after != before)
            assert abs(after - float(lr)) <= 1e-7 * after, (seed, it, after, float(lr))
        assert reductions >= 3 and float(lr) >= 1e-4 * (1 - 1e-6)


def test_graphed_step_recaptures_when_the_learning_rate_changes(monkeypatch):
    """Host logic of ``GraphedStep`` with the CUDA-graph machinery replaced by stand-ins (capture = run the body once,
    replay = nothing): a change of ``param_group['lr']`` -- ``Trainer.update_lr`` does that every epoch,
    trainer_direct.py:122-133 -- must trigger exactly one new capture when the optimiser update is inside the graph,
    and none when it runs eagerly behind the graph or when the learning rate is a device tensor."""
    from ood_dfq_b200 import step

    class FakeGraph:
        replays = 0

        def replay(self):
            FakeGraph.replays += 1

    class FakeCtx:
        def __init__(self, *a):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    class FakeStream:
        def wait_stream(self, other):
            pass

    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", FakeCtx)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", FakeCtx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: FakeStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)

    class Toy:
        """compute / apply split like QATStep, counting body runs."""

        def __init__(self, lr):
            self.w = torch.zeros(3, requires_grad=True)
            self.opt = torch.optim.SGD([self.w], lr=lr, momentum=0.9)
            self.group, self.computes = None, 0

        def compute(self, x):
            self.computes += 1
            return x.sum()

        def apply(self):
            pass

        def __call__(self, x):
            out = self.compute(x)
            self.apply()
            return out

    x = torch.ones(2)
    toy = Toy(0.1)
    g = step.GraphedStep(toy, x, warmup=1)                       # world 1: update inside the graph
    captured = toy.computes
    g(x), g(x)
    assert toy.computes == captured and FakeGraph.replays == 2    # replays only
    toy.opt.param_groups[0]["lr"] = 0.01                          # an epoch boundary
    g(x)
    assert toy.computes == captured + 1                           # one new capture ...
    g(x)
    assert toy.computes == captured + 1                           # ... and replays again
    toy.opt.param_groups[0]["weight_decay"] = 1e-4
    g(x)
    assert toy.computes == captured + 2

    toy2 = Toy(0.1)
    g2 = step.GraphedStep(toy2, x, warmup=1, capture_update=False)   # update runs eagerly behind the graph
    n = toy2.computes
    toy2.opt.param_groups[0]["lr"] = 0.5
    g2(x)
    assert toy2.computes == n

    toy3 = Toy(0.1)
    toy3.opt.param_groups[0]["lr"] = torch.tensor(0.1)             # device-side learning rate: nothing to track
    g3 = step.GraphedStep(toy3, x, warmup=1)
    n = toy3.computes
    toy3.opt.param_groups[0]["lr"].mul_(0.1)
    g3(x)
    assert toy3.computes == n


def test_graphed_step_refuses_to_replay_after_a_mode_change(monkeypatch):
    """freeze / unfreeze / train / eval decide in Python which kernels a forward launches; a captured graph must not be
    replayed across such a change (same stand-in graph machinery as above)."""
    from ood_dfq_b200 import step

    class FakeGraph:
        def replay(self):
            pass

    class FakeCtx:
        def __init__(self, *a):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    class FakeStream:
        def wait_stream(self, other):
            pass

    for name, obj in (("CUDAGraph", FakeGraph), ("graph", FakeCtx), ("Stream", FakeStream), ("stream", FakeCtx)):
        monkeypatch.setattr(torch.cuda, name, obj)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: FakeStream())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    import pytest
    teacher, student, _ = _pair()
    qat = step.QATStep(student, teacher, lr=1e-3, unit_types=(nets.ResUnit,))
    g = step.GraphedStep(qat, torch.randn(2, 3, 32, 32), warmup=0)
    g(torch.randn(2, 3, 32, 32))
    act = next(m for m in student.modules() if isinstance(m, fq_torch.OracleQuantAct))
    act.unfix()
    with pytest.raises(RuntimeError, match="mode flags"):
        g(torch.randn(2, 3, 32, 32))
    act.fix()
    g(torch.randn(2, 3, 32, 32))
    student.train()
    with pytest.raises(RuntimeError, match="mode flags"):
        g(torch.randn(2, 3, 32, 32))


def test_multi_tensor_into_partitions_by_stride_and_keeps_the_sums():
    """``step._multi_tensor_into``: pairs with equal strides go through one foreach call, a pair whose gradient is
    strided differently from its destination (the 1x1 convolution weight: NCHW- vs channels_last-strided, same bytes)
    is moved on its own -- and ``None`` gradients are skipped.  Copy then add gives g1 + g2 exactly."""
    from ood_dfq_b200 import step
    g = torch.Generator().manual_seed(0)
    shapes = [(8,), (4, 3, 3, 3), (6, 5, 1, 1), (7, 2)]
    flat = torch.zeros(sum(int(torch.tensor(s).prod()) for s in shapes))
    dst, off = [], 0
    for s in shapes:
        n = int(torch.tensor(s).prod())
        t = flat[off:off + n].view(s)
        if len(s) == 4:                                     # parameters of a channels_last model: strides (C*H*W, 1, W*C, C)
            t = torch.as_strided(flat, s, (s[1] * s[2] * s[3], 1, s[3] * s[1], s[1]), off)
        dst.append(t)
        off += n
    g1 = [torch.randn(s, generator=g) for s in shapes]
    g2 = [torch.randn(s, generator=g) for s in shapes]
    g1[1] = g1[1].contiguous(memory_format=torch.channels_last)           # same strides as its destination
    g2[1] = g2[1].contiguous(memory_format=torch.channels_last)
    assert g1[2].stride() != dst[2].stride()                              # the odd one out
    g2[3] = None
    calls = []
    real_copy, real_add = torch._foreach_copy_, torch._foreach_add_
    try:
        torch._foreach_copy_ = lambda d, s_: (calls.append(("copy", len(d))), real_copy(d, s_))[1]
        torch._foreach_add_ = lambda d, s_: (calls.append(("add", len(d))), real_add(d, s_))[1]
        step._multi_tensor_into(dst, g1, add=False)
        step._multi_tensor_into(dst, g2, add=True)
    finally:
        torch._foreach_copy_, torch._foreach_add_ = real_copy, real_add
    assert calls == [("copy", 3), ("add", 2)]
    for d, a, b in zip(dst, g1, g2):
        assert torch.equal(d, a if b is None else a + b)


def test_deferred_folds_is_a_gpu_only_context():
    import contextlib
    from ood_dfq_b200 import ops, step
    assert isinstance(step._deferred_folds(torch.zeros(1)), contextlib.nullcontext)
    import pytest
    with pytest.raises(RuntimeError):
        with ops.deferred_folds("cpu"):
            pass
