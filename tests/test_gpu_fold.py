"""Deferred folds of the BatchNorm parameter-gradient reductions (csrc/fold.cu, ``ops.deferred_folds``): inside the block
the reducing backwards only leave notes, one launch folds everything when the block ends -- same bits as the immediate
fold behind every kernel, also when the arena is too small for some of them and across more tensors than one batch holds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bn(c, seed):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(c, generator=g) + 0.5).to(DEV), torch.randn(c, generator=g).to(DEV) * 0.3,
            (torch.randn(c, generator=g) * 0.1).to(DEV), (torch.rand(c, generator=g) + 0.5).to(DEV))


def _work(shapes):
    """A list of closures, each one reducing backward of the library; returns their (dW, dB, ...) outputs."""
    from ood_dfq_b200 import ops
    jobs = []
    for i, shape in enumerate(shapes):
        g = torch.Generator().manual_seed(100 + i)
        c = shape[1]
        x = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
        gy = torch.randn(shape, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
        r = torch.relu(torch.randn(shape, generator=g)).to(DEV).contiguous(memory_format=torch.channels_last)
        ge = torch.randn(shape[:2], generator=g).to(DEV)
        bn1, bn2 = _bn(c, i), _bn(c, 50 + i)
        jobs.append(lambda x=x, gy=gy, bn1=bn1: ops.bn_eval_backward(x, gy, *bn1, 1e-5, relu=True)[1:])
        jobs.append(lambda x=x, gy=gy, r=r, ge=ge, bn1=bn1, bn2=bn2:
                    ops.res_tail_backward(gy, ge, x, r, bn1 + (1e-5,), bn2 + (1e-5,))[2:])
        if shape[2] >= 4 and shape[3] >= 4:
            def pool(x=x, bn1=bn1):
                out, idx, xhat = ops.bn_pool_forward(x, *bn1, 1e-5)
                go = torch.ones_like(out) * 0.5
                return ops.bn_pool_backward(go, idx, xhat, x.shape, *bn1, 1e-5)[1:]
            jobs.append(pool)
    return jobs


def _bits(ts):
    return [t.detach().cpu().numpy().view(np.int32).copy() for t in ts]


@pytest.mark.parametrize("arena_mb", [96, 1])
def test_deferred_folds_equal_immediate_folds(arena_mb):
    from ood_dfq_b200 import _native, ops
    shapes = [(8, 64, 14, 14), (4, 16, 32, 32), (64, 512, 7, 7), (2, 8, 5, 3), (6, 128, 28, 28), (2, 1024, 2, 2)]
    jobs = _work(shapes)
    ref = [_bits(j()) for j in jobs]
    torch.cuda.synchronize()
    lib = _native.load()
    _native.reset_launch_count()
    for j in jobs:
        j()
    immediate_launches = _native.launch_count()
    _native.reset_launch_count()
    with ops.deferred_folds(DEV, arena_mb=arena_mb):
        outs = [j() for j in jobs]
        pending = lib.oodfq_defer_folds_pending()
    deferred_launches = _native.launch_count()
    assert lib.oodfq_defer_folds_pending() == 0
    got = [_bits(o) for o in outs]
    for a, b in zip(ref, got):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    n_reduce = len(jobs)
    if arena_mb >= 96:
        assert pending == n_reduce and deferred_launches == immediate_launches - n_reduce + 1
    else:                                   # a 1 MB arena: some reductions fit, the others fold on the spot
        assert 0 < pending < n_reduce and deferred_launches == immediate_launches - pending + 1


def test_more_deferred_reductions_than_one_batch_holds():
    from ood_dfq_b200 import _native, ops
    jobs = _work([(2, 8, 5, 3)])[:1] * 230           # 230 tiny reductions: three fold launches (96 notes each)
    ref = _bits(jobs[0]())
    _native.reset_launch_count()
    with ops.deferred_folds(DEV):
        outs = [j() for j in jobs]
    assert _native.launch_count() == 230 + 3
    for o in outs:
        for u, v in zip(ref, _bits(o)):
            assert np.array_equal(u, v)


def test_qat_step_with_and_without_deferred_folds_gives_the_same_update(monkeypatch):
    """The whole iteration: weights after two steps are bit-identical whether the folds are deferred or not."""
    import contextlib
    import copy
    from ood_dfq_b200 import fusion, nets, step, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.manual_seed(3)
    teacher = nets.perturb_bn_stats(nets.resnet20_cifar(num_classes=10)).to(DEV).to(memory_format=torch.channels_last)
    student = surgery.quantize_model(copy.deepcopy(teacher).cpu(), 4, 4, namespace=qm).to(DEV).to(memory_format=torch.channels_last)
    g = torch.Generator().manual_seed(4)
    xs = [torch.randn(16, 3, 32, 32, generator=g).to(DEV).contiguous(memory_format=torch.channels_last) for _ in range(2)]
    student.eval()
    with torch.no_grad():
        for x in xs:
            student(x)
    surgery.freeze_model(student, qm)
    for m in (student, teacher):
        fusion.fuse_eval_bn(m, xs[0][:2])
        fusion.fuse_residual_tails(m, xs[0][:2])
    results = []
    for deferred in (True, False):
        s, t = copy.deepcopy(student), copy.deepcopy(teacher)
        if not deferred:
            monkeypatch.setattr(step, "_deferred_folds", lambda like: contextlib.nullcontext())
        q = step.QATStep(s, t, lr=1e-3, unit_types=(nets.ResUnit,) if hasattr(nets, "ResUnit") else ())
        for x in xs:
            q(x)
        results.append([p.detach().clone() for p in s.parameters()])
    for a, b in zip(*results):
        assert torch.equal(a, b)


def test_deferred_statistics_folds_equal_immediate_ones():
    """The BN-input statistics (fp64 sums, ``bn_stats_forward`` / ``bn_eval_stats_forward``) go through the same deferral:
    identical bits, one fold launch for all of them."""
    from ood_dfq_b200 import _native, ops
    shapes = [(8, 64, 14, 14), (4, 16, 32, 32), (16, 512, 7, 7), (2, 8, 5, 3)]
    xs, bns = [], []
    for i, shape in enumerate(shapes):
        g = torch.Generator().manual_seed(200 + i)
        xs.append((torch.randn(shape, generator=g) * 1.3 + 0.2).to(DEV).contiguous(memory_format=torch.channels_last))
        bns.append(_bn(shape[1], 300 + i))

    def run():
        outs = []
        for x, (w, b, rm, rv) in zip(xs, bns):
            outs.append(ops.bn_stats_forward(x, rm))
            sums = torch.empty(2 * x.shape[1], dtype=torch.float64, device=DEV)
            ops.bn_eval_stats_forward(x, w, b, rm, rv, 1e-5, rm, sums, relu=True)
            outs.append(sums)
        return outs

    ref = [o.cpu().numpy().view(np.int64).copy() for o in run()]
    _native.reset_launch_count()
    run()
    immediate = _native.launch_count()
    _native.reset_launch_count()
    with ops.deferred_folds(DEV):
        outs = run()
    assert _native.launch_count() == immediate - 2 * len(shapes) + 1
    for a, o in zip(ref, outs):
        assert np.array_equal(a, o.cpu().numpy().view(np.int64))
