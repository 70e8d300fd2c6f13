"""Drop-in proof on the consumer side: the reference's OWN host code, read from its tree at test time and executed
unmodified against the mirror package.

``main_direct.py`` and ``trainer_direct.py`` cannot be imported here (pyhocon / pytorchcv are absent and
``trainer_direct.py`` has a TabError at :275), but the methods that consume the quantisation API are
self-contained: ``ExperimentDesign.quantize_model`` / ``freeze_model`` / ``unfreeze_model``
(main_direct.py:444-516) and ``Trainer.reduce_minmax`` (trainer_direct.py:368-374).  Their source text is cut out of
the reference files, compiled, and run in a namespace set up the way the reference sets up its own
(``from quantization_utils.quant_modules import *`` after ``ood_dfq_b200.install()``).  Nothing is copied into this
repository; the module skips where ``/root/reference`` does not exist (the GPU box).

Everything here is host logic on CPU tensors: surgery never runs a forward, and ``reduce_minmax`` only touches the
``x_min`` / ``x_max`` buffers (all-reduce in place, then re-assignment), which must keep working on the mirror.
"""
import ast
import copy
import os
import sys
import textwrap
import types

import pytest
import torch
import torch.distributed as dist
from torch import nn

REF = os.environ.get("OODFQ_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REF, "main_direct.py")):
    pytest.skip("reference tree not present (GPU box)", allow_module_level=True)


def reference_namespace():
    """Globals as the reference's scripts see them after their imports (main_direct.py:1-30)."""
    import ood_dfq_b200
    ood_dfq_b200.install()
    ns = {"nn": nn, "torch": torch, "copy": copy, "dist": dist}
    exec("from quantization_utils.quant_modules import *", ns)           # main_direct.py:21, trainer_direct.py:19
    return ns


def methods_of(path, class_name, names):
    """The named methods of one class, compiled from the reference file's own text (ast keeps it verbatim)."""
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    picked = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in picked) == sorted(names)
    module = ast.Module(body=picked, type_ignores=[])
    ns = reference_namespace()
    exec(compile(module, path, "exec"), ns)
    return ns


def lines_of(path, first, last):
    """Lines [first, last] of a file that cannot be parsed as a whole, dedented and compiled on their own."""
    with open(path) as f:
        text = "".join(f.readlines()[first - 1:last])
    ns = reference_namespace()
    exec(compile(textwrap.dedent(text.expandtabs(4)), f"{path}:{first}-{last}", "exec"), ns)
    return ns


@pytest.fixture(scope="module")
def surgery_ns():
    return methods_of(os.path.join(REF, "main_direct.py"), "ExperimentDesign",
                      ["quantize_model", "freeze_model", "unfreeze_model"])


def experiment(ns, qw, qa):
    """A stand-in for ``ExperimentDesign``: only ``self.settings.qw / .qa`` and the three methods are touched."""
    self = types.SimpleNamespace(settings=types.SimpleNamespace(qw=qw, qa=qa))
    for name in ("quantize_model", "freeze_model", "unfreeze_model"):
        setattr(self, name, types.MethodType(ns[name], self))
    return self


def reference_models():
    """The importable model definition of the reference (models.py, 28x28 ResNet-18) plus this repo's carrier nets
    for the pytorchcv-shaped ones."""
    sys.path.insert(0, REF)
    try:
        import models as ref_models
    finally:
        sys.path.remove(REF)
    from ood_dfq_b200 import nets
    torch.manual_seed(0)
    return {"reference models.ResNet18 (28x28)": ref_models.ResNet18(3, 9, img_size=28),
            "resnet20_cifar": nets.resnet20_cifar(num_classes=100),
            "resnet18_imagenet": nets.resnet18_imagenet(num_classes=1000)}


@pytest.mark.parametrize("bits", [(4, 4), (2, 2)])
def test_reference_quantize_model_runs_on_the_mirror(surgery_ns, bits):
    from ood_dfq_b200 import surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    exp = experiment(surgery_ns, *bits)
    for name, net in reference_models().items():
        theirs = exp.quantize_model(net)                                  # reference code, mirror classes
        ours = surgery.quantize_model(copy.deepcopy(net), *bits)          # this repo's restatement of that code
        sd_t, sd_o = theirs.state_dict(), ours.state_dict()
        assert list(sd_t) == list(sd_o), name
        assert all(torch.equal(sd_t[k], sd_o[k]) for k in sd_t), name
        kinds_t = [(n, type(m).__name__) for n, m in theirs.named_modules()]
        kinds_o = [(n, type(m).__name__) for n, m in ours.named_modules()]
        assert kinds_t == kinds_o, name
        convs = [m for m in theirs.modules() if type(m) is qm.Quant_Conv2d]
        acts = [m for m in theirs.modules() if type(m) is qm.QuantAct]
        lins = [m for m in theirs.modules() if type(m) is qm.Quant_Linear]
        assert convs and acts and len(lins) == 1, name
        assert not any(type(m) in (nn.Conv2d, nn.Linear) for m in theirs.modules()), name
        assert all(m.weight_bit == bits[0] for m in convs + lins) and all(m.activation_bit == bits[1] for m in acts)
        # the reference's key shape: "<relu name>.1.x_min" = the QuantAct inside Sequential(ReLU, QuantAct)
        assert any(k.endswith(".1.x_min") for k in sd_t) and any(k.endswith(".1.beta_t") for k in sd_t), name
        assert all(isinstance(p, nn.Parameter) for m in convs for p in [m.weight])


def test_reference_freeze_and_unfreeze_walk_the_mirror(surgery_ns):
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    exp = experiment(surgery_ns, 4, 4)
    for name, net in reference_models().items():
        model = exp.quantize_model(net)
        acts = [m for m in model.modules() if type(m) is qm.QuantAct]
        assert all(m.running_stat for m in acts)
        exp.freeze_model(model)                                           # main_direct.py:486-500, `type(m) == QuantAct`
        assert not any(m.running_stat for m in acts), name
        exp.unfreeze_model(model)
        assert all(m.running_stat for m in acts), name
        assert "Act_min" in repr(acts[0])                                 # quant_modules.py:57-61 (calls .item())


def test_reference_reduce_minmax_runs_on_the_mirror(surgery_ns):
    """trainer_direct.py:368-374 verbatim over a single-process gloo group: in-place all-reduce of the mirror's
    buffers, then re-assignment -- the buffers must stay registered, 1-element, and hold sum / world."""
    ns = lines_of(os.path.join(REF, "trainer_direct.py"), 368, 374)
    exp = experiment(surgery_ns, 4, 4)
    net = reference_models()["resnet20_cifar"]
    model = exp.quantize_model(net)
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    acts = [m for m in model.modules() if isinstance(m, qm.QuantAct)]
    for i, m in enumerate(acts):
        m.x_min.fill_(-0.25 * i)
        m.x_max.fill_(1.0 + i)
    own_group = not dist.is_initialized()
    if own_group:
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        trainer = types.SimpleNamespace(model=types.SimpleNamespace(module=model))      # self.model.module (DDP)
        ns["reduce_minmax"](trainer)
    finally:
        if own_group:
            dist.destroy_process_group()
    for i, m in enumerate(acts):
        assert m.x_min.shape == (1,) and m.x_max.shape == (1,)
        assert m.x_min.item() == -0.25 * i and m.x_max.item() == 1.0 + i
        assert "x_min" in dict(m.named_buffers()) and "x_max" in dict(m.named_buffers())
    assert sum(k.endswith("x_max") for k in model.state_dict()) == len(acts)


def test_reference_dataset_reads_our_shards_and_our_reader_reads_the_same(tmp_path):
    """``direct_dataset`` (main_direct.py:150-209), compiled from the reference file, opens shard files written by
    ``shards.write_shards``; ``shards.load_shards`` returns exactly what it concatenated, and a ``ShardBatches`` fed the
    reference's own ``train_transform`` reproduces its ``__getitem__`` samples under the same seed."""
    pytest.importorskip("torchvision")
    import pickle

    import numpy as np
    import torchvision.transforms as transforms
    from torch.utils.data import Dataset

    from ood_dfq_b200 import shards
    path = os.path.join(REF, "main_direct.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "direct_dataset")
    ns = {"Dataset": Dataset, "transforms": transforms, "pickle": pickle, "np": np, "torch": torch}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)

    rng = np.random.default_rng(0)
    for channels in (3, 1):
        images = rng.standard_normal((22, channels, 28, 28)).astype(np.float32)
        labels = rng.integers(0, 9, 22).astype(np.int64)
        data_prefix, label_prefix = str(tmp_path / f"data{channels}_group"), str(tmp_path / f"labels{channels}_group")
        shards.write_shards(data_prefix, label_prefix, images, labels)
        settings = types.SimpleNamespace(img_size=28, generateDataPath=data_prefix, generateLabelPath=label_prefix)
        logger = types.SimpleNamespace(info=lambda *a, **k: None)
        theirs = ns["direct_dataset"](settings, logger, "pathmnist")
        got_images, got_labels = shards.load_shards(data_prefix, label_prefix)
        assert len(theirs) == 22
        assert np.array_equal(theirs.tmp_data, got_images) and np.array_equal(theirs.tmp_label, got_labels)
        assert np.array_equal(got_images, images) and np.array_equal(got_labels, labels)

        order = shards.rank_indices(22, 0, 1, epoch=0, shuffle=False)
        torch.manual_seed(5)
        want = [theirs[int(i)] for i in order[:8]]                       # the reference's __getitem__, in order
        torch.manual_seed(5)
        mine = shards.ShardBatches(got_images, got_labels, batch=8, shuffle=False, transform=theirs.train_transform,
                                   channels_last=False)
        x, y = next(iter(mine))
        assert x.shape == (8, 3, 28, 28)
        for j, (img, lab) in enumerate(want):
            assert torch.equal(x[j], img) and int(y[j]) == int(lab)


def test_step_losses_equal_the_reference_trainer_source():
    """``Trainer.loss_fn_kd`` / ``loss_fa`` / ``channel_attention`` / ``hook_activation`` / ``hook_fn_forward``
    (trainer_direct.py:308-330, :379-397), compiled from the reference file's text, against the step's host code
    (ood_dfq_b200/step.py) and the oracle's BN hook on the same tensors: bit-identical."""
    from ood_dfq_b200 import step
    from oracle import bns_torch
    path = os.path.join(REF, "trainer_direct.py")
    ns = {}
    for first, last in ((308, 330), (379, 397)):
        ns.update(lines_of(path, first, last))        # `F`, `torch` come from the star import, as in the reference
    g = torch.Generator().manual_seed(8)
    trainer = types.SimpleNamespace(
        settings=types.SimpleNamespace(alpha=20.0, temperature=20.0, lam=1000.0),
        args=types.SimpleNamespace(local_rank="cpu"),
        criterion=nn.CrossEntropyLoss(), KLloss=nn.KLDivLoss(reduction="batchmean"),       # trainer_direct.py:51, :54
        activation=[], activation_teacher=[], mean_list=[], var_list=[], teacher_running_mean=[], teacher_running_var=[])
    trainer.channel_attention = types.MethodType(ns["channel_attention"], trainer)
    for name in ("loss_fn_kd", "loss_fa", "hook_activation", "hook_activation_teacher", "hook_fn_forward"):
        setattr(trainer, name, types.MethodType(ns[name], trainer))

    # KD loss
    s_logits, t_logits = torch.randn(16, 100, generator=g) * 3, torch.randn(16, 100, generator=g) * 3
    labels = torch.randint(0, 100, (16,), generator=g)
    kd_ref, _ = trainer.loss_fn_kd(s_logits, labels, t_logits)
    assert torch.equal(step.kd_loss(s_logits, t_logits, 20.0, 20.0), kd_ref)

    # channel attention + feature-alignment loss over three "units"
    feats_s = [torch.randn(4, c, 6, 6, generator=g) for c in (8, 16, 32)]
    feats_t = [torch.randn(4, c, 6, 6, generator=g) for c in (8, 16, 32)]
    for fs, ft in zip(feats_s, feats_t):
        trainer.hook_activation(None, None, fs)
        trainer.hook_activation_teacher(None, None, ft)
    maps_s, maps_t = [step.channel_attention(f.clone()) for f in feats_s], [step.channel_attention(f.clone()) for f in feats_t]
    assert all(torch.equal(a, b) for a, b in zip(maps_s, trainer.activation))
    assert torch.equal(step.feature_alignment_loss(maps_s, maps_t, 1000.0, "cpu"), trainer.loss_fa())

    # BN-statistics hook
    bn = nn.BatchNorm2d(8).eval()
    bn.running_mean.copy_(torch.randn(8, generator=g))
    x = torch.randn(5, 8, 7, 7, generator=g)
    trainer.hook_fn_forward(bn, (x,), None)
    mean, var = bns_torch.channel_stats(x)
    assert torch.equal(mean, trainer.mean_list[0]) and torch.equal(var, trainer.var_list[0])
    assert trainer.teacher_running_mean[0] is bn.running_mean


def test_qat_step_equals_the_reference_iteration_source():
    """SURVEY row a16.  The student branch of ``Trainer.train`` (trainer_direct.py:500-518) with ``forward`` (:332-340),
    ``loss_fn_kd`` / ``loss_fa`` (:308-330), the feature hooks (:379-386) and ``backward_S`` (:350-356), all compiled
    from the reference file's own text, run one iteration on CPU oracle modules; ``step.QATStep`` runs the same
    iteration on a copy.  Losses and the SGD-updated student weights must agree bit for bit, although the step
    keeps the teacher frozen, prunes the final backward to the student's parameters and owns a flat gradient
    buffer (the documented deviations: none of them may change the update)."""
    from ood_dfq_b200 import nets, step, surgery
    from oracle import fq_torch
    path = os.path.join(REF, "trainer_direct.py")
    ns = {}
    for first, last in ((308, 340), (350, 356), (379, 386)):
        ns.update(lines_of(path, first, last))
    # the loop body itself, as a function of (self, images, labels)
    with open(path) as f:
        body = "".join(f.readlines()[499:518])
    src = "def iteration(self, images, labels):\n" + textwrap.indent(textwrap.dedent(body.expandtabs(4)), "    ") + \
          "    return loss_S, loss_S_perturbed, loss_total\n"
    exec(compile(src, f"{path}:500-518", "exec"), ns)

    torch.manual_seed(3)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=fq_torch)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():                                            # epochs 0-3: calibrate, then freeze (:429, main_direct.py:538)
        for _ in range(2):
            student(torch.randn(4, 3, 32, 32, generator=g))
    surgery.freeze_model(student, fq_torch)
    teacher.eval()
    student.eval()
    images, labels = torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 10, (4,), generator=g)
    hyper = dict(lr=1e-2, momentum=0.9, weight_decay=1e-4)            # a visible update; same rule as lr_S = 1e-5

    # ---- the reference's iteration -------------------------------------------------------------------------------
    ref_student, ref_teacher = copy.deepcopy(student), copy.deepcopy(teacher)
    trainer = types.SimpleNamespace(
        settings=types.SimpleNamespace(alpha=20.0, temperature=20.0, lam=1000.0, eps=0.01),
        args=types.SimpleNamespace(local_rank="cpu"), model=ref_student, model_teacher=ref_teacher,
        criterion=nn.CrossEntropyLoss(), KLloss=nn.KLDivLoss(reduction="batchmean"), activation=[], activation_teacher=[],
        optimizer_S=torch.optim.SGD(params=ref_student.parameters(), nesterov=True, **hyper))      # :59-65
    for name in ("loss_fn_kd", "loss_fa", "forward", "backward_S", "channel_attention", "hook_activation",
                 "hook_activation_teacher", "iteration"):
        setattr(trainer, name, types.MethodType(ns[name], trainer))
    for m in ref_teacher.modules():                                   # :432-440 (ResUnit = the carrier nets' unit class)
        if isinstance(m, nets.ResUnit):
            m.body.register_forward_hook(trainer.hook_activation_teacher)
    for m in ref_student.modules():
        if isinstance(m, nets.ResUnit):
            m.body.register_forward_hook(trainer.hook_activation)
    loss_s, loss_p, loss_total = trainer.iteration(images.clone(), labels)

    # ---- this repository's step -----------------------------------------------------------------------------------
    my_student, my_teacher = copy.deepcopy(student), copy.deepcopy(teacher)
    qat = step.QATStep(my_student, my_teacher, temperature=20.0, alpha=20.0, lam=1000.0, eps=0.01,
                       unit_types=(nets.ResUnit,), **hyper)
    total = qat(images.clone())
    assert torch.equal(total.reshape(-1), loss_total.detach().reshape(-1))
    moved = 0
    for (name, a), (_, b) in zip(my_student.named_parameters(), ref_student.named_parameters()):
        assert torch.equal(a, b), name
        moved += int(not torch.equal(a, dict(student.named_parameters())[name]))
    assert moved > 10                                                 # the update really happened
    assert float(loss_s.detach()) > 0 and float(loss_p.detach()) > 0


def test_distill_step_equals_the_reference_loop_source(monkeypatch):
    """BASELINE config 5 / SURVEY 8(d).  The inner iteration of ``DistillData.getDistilData_hardsample``
    (data_generate/distill_data.py:197-275: the image-size branch, BN hooks' lists, focal cross-entropy, BN-statistics
    terms, Adam step with gradient clipping, ReduceLROnPlateau), compiled from the reference file's text, against
    ``step.DistillStep`` for several iterations on CPU: images, loss and learning rate bit-identical.  The loop
    hard-codes ``.cuda()``; here that call is made the identity (no source change)."""
    import numpy as np
    import torch.nn.functional as F
    from torch import optim

    from ood_dfq_b200 import nets, step
    from oracle import bns_torch
    sys.path.insert(0, REF)
    try:
        from data_generate.distill_data import DistillData
    finally:
        sys.path.remove(REF)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(nn.Module, "cuda", lambda self, *a, **k: self)
    path = os.path.join(REF, "data_generate", "distill_data.py")
    with open(path) as f:
        body = "".join(f.readlines()[196:275])
    src = ("def iteration(self, teacher_model, gaussian_data, labels, gt, beta, gamma, CE_loss, MSE_loss, optimizer, "
           "scheduler, RRC, RHF, i, it):\n" + textwrap.indent(textwrap.dedent(body.expandtabs(4)), "    ") +
           "    return total_loss\n")
    ns = {"torch": torch, "np": np, "F": F, "random": __import__("random"), "print": lambda *a, **k: None}
    exec(compile(src, f"{path}:197-275", "exec"), ns)

    torch.manual_seed(2)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    teacher.img_size = 32                                            # the 28 / 32 branch: no augmentation (:198-204)
    teacher.eval()
    g = torch.Generator().manual_seed(6)
    start = torch.randn(4, 3, 32, 32, generator=g) / 5.0             # :181
    labels = torch.randint(0, 10, (4,), generator=g)

    ref_teacher = copy.deepcopy(teacher)
    dd = DistillData()
    for m in ref_teacher.modules():                                  # :156-158
        if isinstance(m, nn.BatchNorm2d):
            m.register_forward_hook(dd.hook_fn_forward)
    gaussian = start.clone()
    gaussian.requires_grad = True                                    # :182
    optimizer = optim.Adam([gaussian], lr=0.5)                       # :183
    scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, min_lr=1e-4, patience=50)   # :185-188
    ce, mse = nn.CrossEntropyLoss(reduction="none"), nn.MSELoss()    # :149-150

    mine = step.DistillStep(copy.deepcopy(teacher), None, start, labels, lr=0.5, beta=0.1, gamma=0.5)
    mine.stat = bns_torch.StatTap(mine.teacher)
    for it in range(4):
        want = ns["iteration"](dd, ref_teacher, gaussian, labels, labels.numpy(), 0.1, 0.5, ce, mse, optimizer, scheduler,
                               None, None, 0, it)
        got = mine()
        assert torch.equal(got.reshape(-1), want.detach().reshape(-1)), it
        assert torch.equal(mine.images.detach(), gaussian.detach()), it
        assert mine.opt.param_groups[0]["lr"] == optimizer.param_groups[0]["lr"]
    assert not torch.equal(gaussian.detach(), start)


def test_augmented_distill_step_equals_the_reference_loop_source(monkeypatch):
    """The 224-pixel branch of the same loop (distill_data.py:205-227): on every other iteration each image goes
    through torchvision's ``RHF(RRC(gaussian_data[j]))`` before the teacher, and autograd carries the gradient back
    through the resize.  ``DistillStep(augment=...)`` with the oracle's torch augmentation consumes Python's and
    torch's generators the same way: images, loss and learning rate stay bit-identical over iterations of both kinds."""
    import random

    import numpy as np
    import torch.nn.functional as F
    import torchvision.transforms as transforms
    from torch import optim

    from ood_dfq_b200 import nets, step
    from oracle import augment_torch, bns_torch
    sys.path.insert(0, REF)
    try:
        from data_generate.distill_data import DistillData
    finally:
        sys.path.remove(REF)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    path = os.path.join(REF, "data_generate", "distill_data.py")
    with open(path) as f:
        body = "".join(f.readlines()[196:275])
    src = ("def iteration(self, teacher_model, gaussian_data, labels, gt, beta, gamma, CE_loss, MSE_loss, optimizer, "
           "scheduler, RRC, RHF, i, it):\n" + textwrap.indent(textwrap.dedent(body.expandtabs(4)), "    ") +
           "    return total_loss\n")
    ns = {"torch": torch, "np": np, "F": F, "random": random, "print": lambda *a, **k: None}
    exec(compile(src, f"{path}:197-275", "exec"), ns)

    torch.manual_seed(2)
    teacher = nets.resnet20_cifar(num_classes=10)                     # no img_size attribute: the loop's 224 branch (:217-227)
    nets.perturb_bn_stats(teacher)
    teacher.eval()
    side = 32                                                        # the branch is chosen by the missing attribute, not by the size
    g = torch.Generator().manual_seed(6)
    start = torch.randn(3, 3, side, side, generator=g) / 5.0
    labels = torch.randint(0, 10, (3,), generator=g)

    ref_teacher = copy.deepcopy(teacher)
    dd = DistillData()
    for m in ref_teacher.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.register_forward_hook(dd.hook_fn_forward)
    gaussian = start.clone()
    gaussian.requires_grad = True
    optimizer = optim.Adam([gaussian], lr=0.5)
    scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, min_lr=1e-4, patience=50)
    ce, mse = nn.CrossEntropyLoss(reduction="none"), nn.MSELoss()
    rrc = transforms.RandomResizedCrop(size=side, scale=(0.4, 1.0))   # :166-172 with augMargin = 0.4 (:87)
    rhf = transforms.RandomHorizontalFlip()                          # :173

    def oracle_augment(x, boxes, flips):
        return augment_torch.batch(x, range(x.shape[0]), boxes, flips, x.shape[2:], channels=3)

    mine = step.DistillStep(copy.deepcopy(teacher), None, start, labels, lr=0.5, beta=0.1, gamma=0.5,
                            augment=oracle_augment, augment_p=0.5, aug_margin=0.4)
    mine.stat = bns_torch.StatTap(mine.teacher)
    iters, kinds = 6, []
    random.seed(11)
    torch.manual_seed(12)
    want, ref_images = [], []
    for it in range(iters):
        state = random.getstate()
        kinds.append(random.random() < 0.5)                           # peek at the coin the loop is about to toss
        random.setstate(state)
        want.append(ns["iteration"](dd, ref_teacher, gaussian, labels, labels.numpy(), 0.1, 0.5, ce, mse, optimizer,
                                    scheduler, rrc, rhf, 0, it).detach().clone())
        ref_images.append(gaussian.detach().clone())
    assert any(kinds) and not all(kinds)                             # both kinds of iteration were exercised
    random.seed(11)
    torch.manual_seed(12)
    for it in range(iters):
        got = mine()
        assert torch.equal(got.reshape(-1), want[it].reshape(-1)), (it, kinds[it])
        assert torch.equal(mine.images.detach(), ref_images[it]), (it, kinds[it])
    assert mine.opt.param_groups[0]["lr"] == optimizer.param_groups[0]["lr"]


def test_carrier_resnet18_loads_what_the_reference_key_conversion_produces():
    """``convert_state_dict`` (main_direct.py:212-301), compiled from the reference file, maps a torchvision ResNet-18
    checkpoint onto pytorchcv's key names.  The bench's carrier network (ood_dfq_b200/nets.py, shape-faithful stand-in
    for ``ptcv_get_model('resnet18')``) must accept exactly that state dict -- same keys, same shapes -- and then
    computes the same function as the torchvision network."""
    tv = pytest.importorskip("torchvision")
    from collections import OrderedDict

    from ood_dfq_b200 import nets
    path = os.path.join(REF, "main_direct.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "convert_state_dict")
    ns = {"OrderedDict": OrderedDict, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)

    torch.manual_seed(0)
    source = tv.models.resnet18(num_classes=1000)
    for m in source.modules():                                       # non-trivial BN state
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    carrier = nets.resnet18_imagenet(num_classes=1000)
    converted = ns["convert_state_dict"](source.state_dict(), carrier, model_type="standard")
    want = carrier.state_dict()
    assert set(converted) == set(want)
    assert all(converted[k].shape == want[k].shape for k in want)
    carrier.load_state_dict(converted, strict=True)
    source.eval()
    carrier.eval()
    x = torch.randn(2, 3, 224, 224)
    with torch.no_grad():
        assert torch.allclose(carrier(x), source(x), rtol=1e-4, atol=1e-5)


def test_carrier_small_resnet18_is_the_reference_model():
    """``nets.resnet18_small`` (the MedMNIST-shape workload of the bench) against the reference's own ``models.ResNet18``
    (importable): identical state-dict keys and shapes, and the same function once the weights are copied."""
    from ood_dfq_b200 import nets
    sys.path.insert(0, REF)
    try:
        import models as ref_models
    finally:
        sys.path.remove(REF)
    torch.manual_seed(1)
    theirs = ref_models.ResNet18(3, 9, img_size=28)
    mine = nets.resnet18_small(3, 9)
    sd = theirs.state_dict()
    assert list(sd) == list(mine.state_dict())
    assert all(sd[k].shape == v.shape for k, v in mine.state_dict().items())
    mine.load_state_dict(sd, strict=True)
    theirs.eval()
    mine.eval()
    x = torch.randn(3, 3, 28, 28)
    with torch.no_grad():
        assert torch.equal(mine(x), theirs(x))


def test_generator_phase_equals_the_reference_iteration_source():
    """The warm-up branch of ``Trainer.train`` (trainer_direct.py:459-488: generator forward, teacher forward with the
    BN-input hooks, ``CE + 0.1 * BNS``, ``backward_G`` :342-348, then the student's range-tracking forward), compiled
    from the reference file with the reference's own ``Generator_32`` (main_direct.py:52-88), against
    ``step.GeneratorStep`` on CPU oracle modules: losses, the updated generator and every calibrated activation range
    bit for bit."""
    from ood_dfq_b200 import nets, step, surgery
    from oracle import bns_torch, fq_torch
    tpath = os.path.join(REF, "trainer_direct.py")
    ns = {}
    for first, last in ((342, 348), (388, 397)):                     # backward_G, hook_fn_forward
        ns.update(lines_of(tpath, first, last))
    with open(tpath) as f:
        body = "".join(f.readlines()[458:488])
    src = "def iteration(self):\n" + textwrap.indent(textwrap.dedent(body.expandtabs(4)), "    ") + \
          "    return loss_G, loss_one_hot, BNS_loss\n"
    exec(compile(src, f"{tpath}:459-488", "exec"), ns)
    mpath = os.path.join(REF, "main_direct.py")
    with open(mpath) as f:
        tree = ast.parse(f.read(), filename=mpath)
    gen_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Generator_32")
    gns = {"nn": nn, "torch": torch, "Option": None}
    exec(compile(ast.Module(body=[gen_cls], type_ignores=[]), mpath, "exec"), gns)

    settings = types.SimpleNamespace(nClasses=10, latent_dim=100, img_size=32, channels=3)   # cifar10_resnet20.hocon
    torch.manual_seed(7)
    generator = gns["Generator_32"](options=settings)
    teacher = nets.resnet20_cifar(num_classes=10)
    nets.perturb_bn_stats(teacher)
    student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=fq_torch)

    # ---- the reference's iteration -------------------------------------------------------------------------------
    r_gen, r_teacher, r_student = copy.deepcopy(generator), copy.deepcopy(teacher), copy.deepcopy(student)
    r_teacher.eval(), r_student.eval(), r_gen.train()                 # :411-413
    trainer = types.SimpleNamespace(
        settings=settings, args=types.SimpleNamespace(local_rank="cpu"), generator=r_gen, model_teacher=r_teacher,
        model=r_student, criterion=nn.CrossEntropyLoss(), MSE_loss=nn.MSELoss(), mean_list=[], var_list=[],
        teacher_running_mean=[], teacher_running_var=[],
        optimizer_G=torch.optim.Adam(r_gen.parameters(), lr=0.001, betas=(0.5, 0.999)))              # :87-88
    for name in ("backward_G", "hook_fn_forward", "iteration"):
        setattr(trainer, name, types.MethodType(ns[name], trainer))
    for m in r_teacher.modules():                                     # :418-423 (BatchNorm2d here, SyncBatchNorm there)
        if isinstance(m, nn.BatchNorm2d):
            m.register_forward_hook(trainer.hook_fn_forward)
    torch.manual_seed(21)
    want = [tuple(t.detach().clone() for t in trainer.iteration()) for _ in range(3)]

    # ---- this repository's step -----------------------------------------------------------------------------------
    m_gen, m_teacher, m_student = copy.deepcopy(generator), copy.deepcopy(teacher), copy.deepcopy(student)
    gstep = step.GeneratorStep(m_gen, m_teacher, m_student, bns_torch.StatTap(m_teacher), latent_dim=100, n_classes=10,
                               batch=16, lr=0.001, betas=(0.5, 0.999))
    torch.manual_seed(21)
    for it in range(3):
        got = gstep()
        for a, b in zip(got, want[it]):
            assert torch.equal(a.reshape(-1), b.reshape(-1)), it
    for (name, a), (_, b) in zip(m_gen.state_dict().items(), r_gen.state_dict().items()):
        assert torch.equal(a, b), name
    acts_m = [m for m in m_student.modules() if isinstance(m, fq_torch.OracleQuantAct)]
    acts_r = [m for m in r_student.modules() if isinstance(m, fq_torch.OracleQuantAct)]
    assert len(acts_m) == len(acts_r) > 10
    for a, b in zip(acts_m, acts_r):
        assert torch.equal(a.x_min, b.x_min) and torch.equal(a.x_max, b.x_max) and torch.equal(a.beta_t, b.beta_t)
        assert a.x_max.item() > 0 and a.beta_t.item() < 1            # the ranges really were calibrated


def test_fusion_passes_recognise_the_reference_model_classes():
    """The passes find the BatchNorm / residual-unit layout of the reference's OWN ``models.ResNet18`` (BasicBlock:
    conv1..bn2 / shortcut / relu2, models.py:21-47), swap classes without touching keys, and -- off the GPU, where the
    fused modules run their class's own forward -- leave the function unchanged."""
    from ood_dfq_b200 import fusion, nets
    sys.path.insert(0, REF)
    try:
        import models as ref_models
    finally:
        sys.path.remove(REF)
    torch.manual_seed(5)
    model = ref_models.ResNet18(3, 9, img_size=28).eval()
    nets.perturb_bn_stats(model)
    x = torch.randn(2, 3, 28, 28)
    with torch.no_grad():
        ref = model(x)
    keys = list(model.state_dict())
    units_before = [type(m) for m in model.modules() if isinstance(m, ref_models.BasicBlock)]
    fusion.fuse_eval_bn(model, x)
    assert fusion.fuse_residual_tails(model, x) == len(units_before) == 8
    assert all(isinstance(m, fusion.FusedEvalBN) for m in model.modules() if isinstance(m, nn.BatchNorm2d))
    assert sum(isinstance(m, ref_models.BasicBlock) for m in model.modules()) == 8          # still BasicBlocks
    with torch.no_grad():
        assert torch.equal(model(x), ref)
    assert list(model.state_dict()) == keys


def test_checkpoints_move_between_the_reference_classes_and_the_mirror(surgery_ns):
    """A model quantised with the reference's OWN classes and one quantised with the mirror's have the same state dict
    (keys, shapes, values), load each other's checkpoints strictly, survive ``convert_sync_batchnorm``
    (main_direct.py:483) and wrap into DDP (:484) -- on CPU, gloo, no forward."""
    import importlib
    pkg = types.ModuleType("_live_reference_qu2")
    pkg.__path__ = [os.path.join(REF, "quantization_utils")]
    sys.modules["_live_reference_qu2"] = pkg
    ref_qm = importlib.import_module("_live_reference_qu2.quant_modules")

    # the reference's quantize_model, once bound to its own classes and once to the mirror's
    path = os.path.join(REF, "main_direct.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExperimentDesign")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "quantize_model")
    ns_ref = {"nn": nn, "torch": torch, "copy": copy}
    ns_ref.update({k: getattr(ref_qm, k) for k in ("Quant_Conv2d", "Quant_Linear", "QuantAct")})
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns_ref)
    exp_ref = types.SimpleNamespace(settings=types.SimpleNamespace(qw=4, qa=4))
    exp_ref.quantize_model = types.MethodType(ns_ref["quantize_model"], exp_ref)
    exp_mine = experiment(surgery_ns, 4, 4)

    net = reference_models()["reference models.ResNet18 (28x28)"]
    theirs, mine = exp_ref.quantize_model(net), exp_mine.quantize_model(net)
    assert type(next(m for m in theirs.modules() if "QuantAct" in type(m).__name__)).__module__.startswith("_live_reference")
    sd_t, sd_m = theirs.state_dict(), mine.state_dict()
    assert list(sd_t) == list(sd_m) and all(sd_t[k].shape == sd_m[k].shape and torch.equal(sd_t[k], sd_m[k]) for k in sd_t)
    for m in theirs.modules():                              # a "trained" checkpoint: non-default ranges
        if hasattr(m, "x_min"):
            m.x_min.fill_(-0.125)
            m.x_max.fill_(3.5)
            m.beta_t.fill_(0.25)
    mine.load_state_dict(theirs.state_dict(), strict=True)
    assert all(m.x_max.item() == 3.5 and m.beta_t.item() == 0.25 for m in mine.modules() if hasattr(m, "x_min"))
    theirs.load_state_dict(mine.state_dict(), strict=True)

    sync = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(mine))
    assert list(sync.state_dict()) == list(sd_m)
    assert sum(isinstance(m, nn.SyncBatchNorm) for m in sync.modules()) == sum(isinstance(m, nn.BatchNorm2d) for m in mine.modules())

    own_group = not dist.is_initialized()
    if own_group:
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        ddp = torch.nn.parallel.DistributedDataParallel(mine, broadcast_buffers=False)
        assert [k[len("module."):] for k in ddp.state_dict()] == list(sd_m)
        assert sum(p.numel() for p in ddp.parameters()) == sum(p.numel() for p in net.parameters())
    finally:
        if own_group:
            dist.destroy_process_group()


def test_reference_training_code_drives_the_mirror_modules(surgery_ns):
    """The whole consumer chain at once: the reference's own ``quantize_model`` builds a student from the MIRROR
    classes, its own warm-up iteration (:459-488) calibrates it, its own ``freeze_model`` freezes it and its own QAT
    iteration (:500-518) trains it for several steps -- with the mirror's kernel launches swapped for oracle arithmetic
    (tests/cpu_ops_shim.py), so that everything else is the product's host code: autograd Functions, in-place range
    updates, the WeightBank across ``optimizer_S.step()``.  A student built from the reference's OWN classes goes
    through the same code; ranges, losses and weights must stay bit-identical all the way."""
    import importlib

    import cpu_ops_shim
    from ood_dfq_b200 import nets
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    pkg = types.ModuleType("_live_reference_qu4")
    pkg.__path__ = [os.path.join(REF, "quantization_utils")]
    sys.modules["_live_reference_qu4"] = pkg
    ref_qm = importlib.import_module("_live_reference_qu4.quant_modules")

    tpath, mpath = os.path.join(REF, "trainer_direct.py"), os.path.join(REF, "main_direct.py")
    with open(mpath) as f:
        tree = ast.parse(f.read(), filename=mpath)
    exp_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExperimentDesign")
    wanted = [n for n in exp_cls.body if isinstance(n, ast.FunctionDef) and n.name in ("quantize_model", "freeze_model")]
    gen_cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Generator_32")
    with open(tpath) as f:
        lines = f.readlines()
    warm = "def warm_up(self):\n" + textwrap.indent(textwrap.dedent("".join(lines[458:488]).expandtabs(4)), "    ") + \
           "    return loss_G\n"
    train = "def train_step(self, images, labels):\n" + \
            textwrap.indent(textwrap.dedent("".join(lines[499:518]).expandtabs(4)), "    ") + "    return loss_total\n"

    def side(classes):
        """One complete trainer stand-in whose namespace binds the quantisation names to ``classes``."""
        ns = {"nn": nn, "torch": torch, "copy": copy, "Option": None}
        exec("from torch.autograd import Variable\nimport torch.nn.functional as F", ns)
        ns.update({k: getattr(classes, k) for k in ("Quant_Conv2d", "Quant_Linear", "QuantAct")})
        exec(compile(ast.Module(body=wanted + [gen_cls], type_ignores=[]), mpath, "exec"), ns)
        for first, last in ((308, 340), (342, 348), (350, 356), (379, 397)):
            exec(compile(textwrap.dedent("".join(lines[first - 1:last]).expandtabs(4)), tpath, "exec"), ns)
        exec(compile(warm, tpath, "exec"), ns)
        exec(compile(train, tpath, "exec"), ns)
        settings = types.SimpleNamespace(qw=4, qa=4, nClasses=10, latent_dim=100, img_size=32, channels=3, alpha=20.0,
                                         temperature=20.0, lam=1000.0, eps=0.01)
        torch.manual_seed(9)
        teacher = nets.resnet20_cifar(num_classes=10)
        nets.perturb_bn_stats(teacher)
        t = types.SimpleNamespace(settings=settings, args=types.SimpleNamespace(local_rank="cpu"), model_teacher=teacher.eval(),
                                  generator=ns["Generator_32"](options=settings).train(), criterion=nn.CrossEntropyLoss(),
                                  MSE_loss=nn.MSELoss(), KLloss=nn.KLDivLoss(reduction="batchmean"), mean_list=[], var_list=[],
                                  teacher_running_mean=[], teacher_running_var=[], activation=[], activation_teacher=[])
        for name in ("quantize_model", "freeze_model", "loss_fn_kd", "loss_fa", "forward", "backward_G", "backward_S",
                     "channel_attention", "hook_activation", "hook_activation_teacher", "hook_fn_forward", "warm_up",
                     "train_step"):
            setattr(t, name, types.MethodType(ns[name], t))
        t.model = t.quantize_model(copy.deepcopy(teacher)).eval()
        t.optimizer_G = torch.optim.Adam(t.generator.parameters(), lr=1e-3, betas=(0.5, 0.999))
        t.optimizer_S = torch.optim.SGD(t.model.parameters(), lr=1e-5, momentum=0.9, weight_decay=1e-4, nesterov=True)   # lr_S of cifar10_resnet20.hocon
        handles = [m.register_forward_hook(t.hook_fn_forward) for m in teacher.modules() if isinstance(m, nn.BatchNorm2d)]
        return t, handles

    def run(t, handles):
        out = []
        torch.manual_seed(33)
        for _ in range(2):                                           # epochs 0-3: generator + range calibration
            out.append(t.warm_up().detach().clone())
        for h in handles:                                            # epoch 4 (:425-440): hooks off, feature taps on
            h.remove()
        t.freeze_model(t.model)
        for m in t.model_teacher.modules():
            if isinstance(m, nets.ResUnit):
                m.body.register_forward_hook(t.hook_activation_teacher)
        for m in t.model.modules():
            if isinstance(m, nets.ResUnit):
                m.body.register_forward_hook(t.hook_activation)
        g = torch.Generator().manual_seed(34)
        for _ in range(3):
            images, labels = torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 10, (4,), generator=g)
            out.append(t.train_step(images, labels).detach().clone())
        return out

    ref_side, ref_handles = side(ref_qm)
    want = run(ref_side, ref_handles)
    with cpu_ops_shim.installed():
        my_side, my_handles = side(qm)
        assert type(next(m for m in my_side.model.modules() if type(m).__name__ == "QuantAct")) is qm.QuantAct
        got = run(my_side, my_handles)
        for i, (a, b) in enumerate(zip(got, want)):
            assert torch.equal(a.reshape(-1), b.reshape(-1)), i
        sd_a, sd_b = my_side.model.state_dict(), ref_side.model.state_dict()
        assert list(sd_a) == list(sd_b)
        for k in sd_a:
            assert torch.equal(sd_a[k].reshape(-1), sd_b[k].reshape(-1)), k
        assert any(k.endswith("x_max") and sd_a[k].item() > 0 for k in sd_a)
        assert all(torch.isfinite(v).all() for v in sd_a.values())


def test_fusion_passes_on_the_reference_model_quantised_by_the_reference_code(surgery_ns):
    """``models.ResNet18`` (in-place shortcut add, in-place ReLUs) -> the reference's ``quantize_model`` with mirror
    classes -> calibrate -> the reference's ``freeze_model`` -> all fusion passes.  Nine BN-fed tails absorbed, the eight
    tails behind the in-place adds fused as ReLU+QuantAct pairs (NOT swallowed by bn2), eight units recognised; results
    and keys unchanged off the GPU."""
    import cpu_ops_shim
    from ood_dfq_b200 import fusion, nets
    sys.path.insert(0, REF)
    try:
        import models as ref_models
    finally:
        sys.path.remove(REF)
    torch.manual_seed(6)
    net = ref_models.ResNet18(3, 9, img_size=28)
    nets.perturb_bn_stats(net)
    exp = experiment(surgery_ns, 4, 4)
    x = torch.randn(2, 3, 28, 28)
    with cpu_ops_shim.installed():
        student = exp.quantize_model(net).eval()
        with torch.no_grad():
            student(x)
        exp.freeze_model(student)
        with torch.no_grad():
            ref = student(x)
        keys = list(student.state_dict())
        fusion.fuse_eval_bn(student, x)
        units = fusion.fuse_residual_tails(student, x)
        count = lambda cls: sum(type(m) is cls for m in student.modules())
        assert (count(fusion.AbsorbedTail), count(fusion.FusedReLUQuant), units) == (9, 8, 8)
        with torch.no_grad():
            assert torch.equal(student(x), ref)
        assert list(student.state_dict()) == keys
        # the reference's attribute walk (dir() / getattr over every module, exact-type checks) still reaches every
        # QuantAct through the class-swapped containers
        acts = [m for m in student.modules() if type(m).__name__ == "QuantAct"]
        exp.unfreeze_model(student)
        assert len(acts) == 17 and all(m.running_stat for m in acts)
        exp.freeze_model(student)
        assert not any(m.running_stat for m in acts)


def test_reference_bsdc_correction_runs_on_a_fused_mirror_student(surgery_ns):
    """``Trainer._init_bn_tracking`` / ``_create_bn_stat_hook`` / ``apply_bsdc_correction`` (trainer_direct.py:135-307,
    the BN-statistic delta correction: forward hooks on EVERY BatchNorm of teacher and student, both nets switched to
    ``train()`` under ``no_grad``, running statistics rewritten through ``.data.copy_``), compiled from the reference
    file -- whose lines 275-277 are the space-indented ones that make the file as a whole a TabError; tabs are expanded
    to 8 columns for this cut, nothing else is touched -- against (a) a student built from the reference's own
    classes and (b) a mirror-built student AFTER all fusion passes (kernel launches swapped for oracle arithmetic).
    The fused modules must step aside (training mode, hooked BatchNorms) so that every statistic BSDC collects and
    every running statistic it writes is bit-identical."""
    import importlib
    from collections import OrderedDict

    import cpu_ops_shim
    from ood_dfq_b200 import fusion, nets
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    pkg = types.ModuleType("_live_reference_qu5")
    pkg.__path__ = [os.path.join(REF, "quantization_utils")]
    sys.modules["_live_reference_qu5"] = pkg
    ref_qm = importlib.import_module("_live_reference_qu5.quant_modules")
    sys.path.insert(0, REF)
    try:
        import models as ref_models
    finally:
        sys.path.remove(REF)
    tpath = os.path.join(REF, "trainer_direct.py")
    with open(tpath) as f:
        text = "".join(f.readlines()[134:307])
    ns = {"torch": torch, "nn": nn, "dist": dist, "OrderedDict": OrderedDict}
    exec(compile(textwrap.dedent(text.expandtabs(8)), f"{tpath}:135-307", "exec"), ns)

    class Wrapped(nn.Module):                                         # stands in for DistributedDataParallel (.module)
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, x):
            return self.module(x)

    def trainer_for(classes, fuse):
        torch.manual_seed(3)
        # the reference's own 28x28 ResNet-18: BSDC pairs teacher and student BatchNorms BY NAME (:137-151), which only
        # works where quantize_model keeps the names -- index-named Sequentials, as in models.py
        teacher = ref_models.ResNet18(3, 9, img_size=28)
        nets.perturb_bn_stats(teacher)
        path = os.path.join(REF, "main_direct.py")
        with open(path) as f:
            tree = ast.parse(f.read(), filename=path)
        cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExperimentDesign")
        fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "quantize_model")
        qns = {"nn": nn, "torch": torch, "copy": copy}
        qns.update({k: getattr(classes, k) for k in ("Quant_Conv2d", "Quant_Linear", "QuantAct")})
        exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), qns)
        exp = types.SimpleNamespace(settings=types.SimpleNamespace(qw=4, qa=4))
        exp.quantize_model = types.MethodType(qns["quantize_model"], exp)
        student = exp.quantize_model(copy.deepcopy(teacher)).eval()
        teacher.eval()
        x = torch.randn(2, 3, 28, 28, generator=torch.Generator().manual_seed(4))
        with torch.no_grad():
            student(x)                                                # calibrate once, then freeze
        for m in student.modules():
            if type(m).__name__ == "QuantAct":
                m.fix()
        if fuse:
            for net in (student, teacher):
                fusion.fuse_eval_bn(net, x)
                assert fusion.fuse_residual_tails(net, x) == 8
        t = types.SimpleNamespace(
            model=Wrapped(student).eval(), model_teacher=Wrapped(teacher).eval(), logger=None, args=types.SimpleNamespace(local_rank="cpu"),
            settings=types.SimpleNamespace(nEpochs=150), bsdc_num_batches=None, bsdc_correction_applied=False,
            bn_layer_names=[], teacher_bn_layers=[], student_bn_layers=[], teacher_bn_source_stats=[], bsdc_delta_means=[],
            bsdc_delta_vars=[], bsdc_teacher_ood_stats=[], bsdc_student_ood_stats=[])
        for name in ("_init_bn_tracking", "_create_bn_stat_hook", "apply_bsdc_correction"):
            setattr(t, name, types.MethodType(ns[name], t))
        t._init_bn_tracking()
        return t, student, teacher

    g = torch.Generator().manual_seed(5)
    loader = [(torch.randn(4, 3, 28, 28, generator=g) * 1.5 + 0.2, torch.zeros(4, dtype=torch.long)) for _ in range(3)]

    ref_t, ref_student, ref_teacher = trainer_for(ref_qm, fuse=False)
    ref_t.apply_bsdc_correction(loader, epoch=149)
    with cpu_ops_shim.installed():
        my_t, my_student, my_teacher = trainer_for(qm, fuse=True)
        assert any(isinstance(m, fusion._FusedUnitMixin) for m in my_student.modules())
        before = {k: v.clone() for k, v in my_student.state_dict().items()}
        my_t.apply_bsdc_correction(loader, epoch=149)
    assert len(ref_t.student_bn_layers) == len(my_t.student_bn_layers) == 20 and my_t.bsdc_correction_applied
    sd_ref, sd_my = ref_student.state_dict(), my_student.state_dict()
    assert list(sd_ref) == list(sd_my)
    for k in sd_ref:
        assert torch.equal(sd_ref[k].reshape(-1), sd_my[k].reshape(-1)), k
    for a, b in zip(ref_t.bsdc_delta_means + ref_t.bsdc_delta_vars, my_t.bsdc_delta_means + my_t.bsdc_delta_vars):
        assert torch.equal(a, b)
    assert any(k.endswith("running_mean") and not torch.equal(before[k], sd_my[k]) for k in sd_my)   # BSDC did write
    for (k, a), (_, b) in zip(ref_teacher.state_dict().items(), my_teacher.state_dict().items()):
        assert torch.equal(a, b), k
    assert not my_student.training and not my_teacher.training


@pytest.mark.parametrize("net_name,side", [("resnet18_imagenet", 224), ("resnet20_cifar", 32), ("resnet18_small", 28)])
def test_reference_freeze_walk_reaches_every_quantact_of_a_fused_model(surgery_ns, net_name, side):
    """``freeze_model`` / ``unfreeze_model`` run after (and before) EVERY epoch (main_direct.py:533-538) on whatever the
    fusion passes left behind: the walk must still reach each QuantAct -- also through a DDP-style wrapper."""
    import cpu_ops_shim
    from ood_dfq_b200 import fusion, nets
    torch.manual_seed(1)
    base = nets.resnet18_small(3, 9) if net_name == "resnet18_small" else getattr(nets, net_name)(num_classes=10)
    exp = experiment(surgery_ns, 4, 4)
    x = torch.randn(2, 3, side, side)
    with cpu_ops_shim.installed():
        student = exp.quantize_model(base).eval()
        with torch.no_grad():
            student(x)
        exp.freeze_model(student)
        fusion.fuse_eval_bn(student, x)
        fusion.fuse_residual_tails(student, x)
        fusion.space_to_depth_stem(student, x)
    acts = [m for m in student.modules() if type(m).__name__ == "QuantAct"]
    wrapped = nn.Sequential()                                         # any container with a `.module`-like child
    wrapped.module = student
    for target in (student, wrapped):
        exp.unfreeze_model(target)
        assert len(acts) >= 17 and all(m.running_stat for m in acts), net_name
        exp.freeze_model(target)
        assert not any(m.running_stat for m in acts), net_name


def test_reference_option_class_runs_on_our_hocon_reader():
    """``Option.__init__`` (options.py:12-71), compiled from the reference file with ``ConfigFactory.parse_file`` served by
    ``ood_dfq_b200.hocon.load`` (pyhocon is not installed): every one of the reference's config files must give the
    reference's own option object every key it asks for, and ``QuantSettings`` must agree with it on the fields that
    reach the quantisation path."""
    import glob
    import shutil
    import uuid

    from ood_dfq_b200 import hocon
    path = os.path.join(REF, "options.py")
    with open(path) as f:
        tree = ast.parse(f.read(), filename=path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Option")

    class Conf(dict):                                                 # pyhocon's ConfigTree: mapping with .get
        pass

    ns = {"os": os, "shutil": shutil, "uuid": uuid, "NetOption": object,
          "ConfigFactory": types.SimpleNamespace(parse_file=lambda p: Conf(hocon.load(p)))}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    files = sorted(glob.glob(os.path.join(REF, "config", "*.hocon")))
    assert len(files) == 15
    for f in files:
        opt = ns["Option"](f)
        mine = hocon.QuantSettings.from_file(f)
        for name in ("model_name", "dataset", "batchSize", "nClasses", "img_size", "channels", "qw", "qa", "temperature",
                     "alpha", "lr_S", "momentum", "weightDecay", "lam", "eps"):
            assert getattr(mine, name) == getattr(opt, name), (os.path.basename(f), name)
        assert isinstance(opt.step_S, list) and opt.lrPolicy_S == "multi_step" and opt.nEpochs > 0
        assert opt.latent_dim > 0 and 0 < opt.b1 < 1 and opt.bsdc_start_epoch == opt.nEpochs - 1
