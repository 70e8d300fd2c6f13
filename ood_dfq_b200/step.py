"""The steady-state QAT iteration that drives the quantisation path, restated as runnable host code.

This is the CALLER of the hot path, not the path: the reference's version lives in
``Trainer.train`` (trainer_direct.py:490-518, helpers :308-340, :350-356, :379-386), a file
that cannot even be compiled (TabError at :275) and needs pytorchcv.  Per iteration:

    teacher(images)  [grad to images]  ->  student(images)  ->  KD + feature-alignment loss
    -> sign of d loss / d images  ->  images + eps * sign  ->  teacher (no grad) + student again
    -> backward of both losses  ->  SGD(nesterov) step

so the quantisation path runs in two student forwards and is traversed by three backward
sweeps.  The same code drives the CUDA mirror on a GPU and the CPU oracle modules (the
``--impl reference`` arm of bench.py), only the module classes differ.

Deviation from the reference, kept on purpose and identical in both arms: the teacher's
parameters do not require grad (the reference leaves them trainable and lets DDP all-reduce
gradients nobody uses); the student update is unaffected.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


def kd_loss(student_logits, teacher_logits, temperature: float, alpha: float):
    """``KLDiv(log_softmax(s/T), softmax(t/T), 'batchmean') * alpha * T^2`` (trainer_direct.py:308-323)."""
    a = F.log_softmax(student_logits / temperature, dim=1)
    b = F.softmax(teacher_logits / temperature, dim=1)
    return F.kl_div(a, b, reduction="batchmean") * (alpha * temperature * temperature)


def channel_attention(x):
    """``F.normalize(x.pow(2).mean([2,3]))`` exactly as the reference writes it (trainer_direct.py:382-383)."""
    return F.normalize(x.pow(2).mean([2, 3]).view(x.size(0), -1))


class _ChannelEnergy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from . import ops
        ctx.save_for_backward(x)
        return ops.channel_energy_forward(x)

    @staticmethod
    def backward(ctx, grad_e):
        from . import ops
        (x,) = ctx.saved_tensors
        return ops.channel_energy_backward(x, grad_e)


def channel_attention_fused(x):
    """Same quantity with the per-(image, channel) mean of squares as one sm_100a kernel forward and one
    backward (the clone / pow / mean chain and its tape are 20 + 24 B/elem in eager PyTorch, 4 + 8 here)."""
    return F.normalize(_ChannelEnergy.apply(x))


class FeatureTap:
    """Forward hooks collecting channel attention of the residual bodies (trainer_direct.py:432-440).

    ``fused``: use the kernels (CUDA tensors) -- ``maps`` then holds the RAW per-(image, channel) energies
    ``x.pow(2).mean([2,3])`` and the normalisation happens inside the fused loss (``feature_alignment_loss(...,
    raw=True)``); otherwise the reference's expression on a clone of the output, normalised as there.
    """

    def __init__(self, model: nn.Module, unit_types: tuple, fused=False):
        self.maps, self.fused = [], fused
        self.handles = [m.body.register_forward_hook(self._hook) for m in model.modules()
                        if isinstance(m, unit_types) and hasattr(m, "body")]

    def _hook(self, module, inputs, output):
        self.maps.append(_ChannelEnergy.apply(output) if self.fused else channel_attention(output.clone()))

    def clear(self):
        self.maps.clear()


class _FusedFALoss(torch.autograd.Function):
    """``lam * sum_l mean((normalize(Es_l) - normalize(Et_l))^2)`` and its gradient as one launch each
    (csrc/fa_loss.cu) instead of ~20 element-wise / reduction launches per unit and pass."""

    @staticmethod
    def forward(ctx, lam, n_units, *energies):
        from . import ops
        es, et = list(energies[:n_units]), list(energies[n_units:])
        ctx.save_for_backward(*energies)
        ctx.lam, ctx.n_units = lam, n_units
        return ops.fa_loss_forward(es, et, lam)

    @staticmethod
    def backward(ctx, grad_loss):
        from . import ops
        L = ctx.n_units
        es, et = list(ctx.saved_tensors[:L]), list(ctx.saved_tensors[L:])
        need = ctx.needs_input_grad[2:]
        gs, gt = ops.fa_loss_backward(es, et, ctx.lam, grad_loss, want_student=any(need[:L]), want_teacher=any(need[L:]))
        return (None, None) + tuple(g if n else None for g, n in zip(gs + gt, need))


def feature_alignment_loss(student_maps, teacher_maps, lam: float, device, raw=False):
    """``lam * sum_l mean((A_s - A_t)^2)`` (trainer_direct.py:325-330).  ``raw``: the maps are un-normalised
    energies (``FeatureTap(fused=True)``) and the whole expression, normalisation included, runs as one kernel."""
    if raw and len(student_maps) > 0:
        return _FusedFALoss.apply(float(lam), len(student_maps), *student_maps, *teacher_maps)
    fa = torch.zeros(1, device=device)
    for s, t in zip(student_maps, teacher_maps):
        fa = fa + (s - t).pow(2).mean()
    return lam * fa


class FlatGrads:
    """All gradients of a model as views into ONE buffer: one memset, one NCCL all-reduce per step.

    Replaces DDP's bucketed reducer for this step (main_direct.py:484): the student is
    traversed by ``autograd.grad`` and two forwards before its single ``backward``, a
    pattern DDP only tolerates, and the whole payload (46.7 MB for ResNet-18) is one
    NVLink-speed collective anyway.
    """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            # same strides as the parameter (NCHW or channels_last), so the optimiser's element-wise
            # updates stay on the vectorised same-layout path
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            p.grad = (torch.as_strided(self.flat, p.shape, p.stride(), off) if dense
                      else self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()

    def attached(self):
        """True while every parameter's .grad still aliases the flat buffer (Module.to() would break that)."""
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * self.flat.element_size()
        return all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params)

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


def _deferred_folds(like):
    """``ops.deferred_folds`` on the device of ``like`` when that is a GPU and the model uses this package's kernels,
    else a no-op context (the CPU arm drives oracle / reference modules)."""
    import contextlib
    import os
    if not like.is_cuda or os.environ.get("OODFQ_NO_DEFERRED_FOLDS"):        # (the switch is for A/B measurements)
        return contextlib.nullcontext()
    from . import ops
    return ops.deferred_folds(like.device)


def _multi_tensor_into(dst, src, add=False):
    """``dst[i].copy_(src[i])`` (or ``+=``) for every ``src[i]`` that is not None, as few launches as torch allows.

    ``torch._foreach_*`` takes its multi-tensor path only when EVERY pair of a call has equal strides; one 1x1
    convolution weight whose gradient arrives NCHW-strided while the parameter is channels_last-strided (the same
    bytes) sends all 62 tensors down the per-tensor path.  Pairs with equal strides therefore go in one call, the few
    others one by one."""
    fast_d, fast_s = [], []
    for d, g in zip(dst, src):
        if g is None:
            continue
        if d.stride() == g.stride() and d.shape == g.shape and d.dtype == g.dtype and d.device == g.device:
            fast_d.append(d)
            fast_s.append(g)
        elif add:
            d.add_(g)
        else:
            d.copy_(g)
    if fast_d:
        if add:
            torch._foreach_add_(fast_d, fast_s)
        else:
            torch._foreach_copy_(fast_d, fast_s)


def _weight_bank(model):
    """The ``WeightBank`` of this package's weight-quantising modules when ``model`` is built from them, else None
    (the CPU arm drives oracle / reference modules, which re-quantise on every forward)."""
    from .quantization_utils import quant_modules as qm
    return qm.WeightBank if any(isinstance(m, qm._WeightQuantBase) for m in model.modules()) else None


def _drop_stem_cache():
    """The space-to-depth stem keeps the re-laid-out batch of the last call so that teacher and student share it
    within an iteration (fusion._S2DCache); nothing may outlive the iteration (a batch and its copy, 300 MB)."""
    from .fusion import _S2DCache
    _S2DCache.clear()


class QATStep:
    """One data-free QAT iteration (see module docstring).  ``__call__`` returns the detached losses."""

    def __init__(self, student, teacher, lr=1e-6, momentum=0.9, weight_decay=1e-4, temperature=20.0,
                 alpha=20.0, lam=1000.0, eps=0.01, unit_types: tuple = (), group=None, perturb=True,
                 fused_attention=None, prune_backward=True, fused_optimizer=True):
        self.student, self.teacher = student, teacher
        self.T, self.alpha, self.lam, self.eps = temperature, alpha, lam, eps
        self.group, self.perturb = group, perturb
        # False: never exchange gradients, even with an initialised process group (a single-process run of the GLOBAL
        # batch inside a multi-rank job: the oracle of the data-parallel parity checks)
        self.exchange = True
        # The final backward only has to deliver the student's parameter gradients.  ``loss.backward()`` as the
        # reference writes it (trainer_direct.py:350-356) also walks the teacher's graph and the student's stem
        # dgrad down to ``images.grad``, which nothing reads after the sign perturbation; with
        # ``backward(inputs=params)`` autograd prunes those branches.  The update is bit-identical
        # (tests/test_dist_gloo.py, tests/test_gpu_fused.py); ``prune_backward=False`` restores the full sweep.
        self.prune_backward = prune_backward
        for p in teacher.parameters():
            p.requires_grad_(False)
        self.grads = FlatGrads(student.parameters())
        # torch's multi-tensor ("fused") SGD where the parameters live on a GPU: the update of all 62 tensors is one
        # launch per 2^n-tensor chunk instead of ~350 element-wise launches (0.9 ms of a 35 ms step); same formula as
        # the reference's torch.optim.SGD(nesterov=True) (trainer_direct.py:59-65), stock PyTorch either way
        on_gpu = self.grads.params[0].is_cuda
        self.opt = torch.optim.SGD(self.grads.params, lr=lr, momentum=momentum, weight_decay=weight_decay,
                                   nesterov=True, **({"fused": True} if on_gpu and fused_optimizer else {}))
        if fused_attention is None:             # the kernel where the model lives on a GPU; the CPU arm keeps torch
            fused_attention = next(student.parameters()).is_cuda
        self.tap_s = FeatureTap(student, unit_types, fused_attention) if unit_types else None
        self.tap_t = FeatureTap(teacher, unit_types, fused_attention) if unit_types else None
        student.eval()       # trainer_direct.py:411-412: both nets use BN running statistics
        teacher.eval()

    def _losses(self, images, teacher_logits):
        out = self.student(images)
        kl = kd_loss(out, teacher_logits, self.T, self.alpha)
        if self.tap_s is not None:
            fa = feature_alignment_loss(self.tap_s.maps, self.tap_t.maps, self.lam, images.device,
                                        raw=self.tap_s.fused and self.tap_t.fused)
        else:
            fa = torch.zeros(1, device=images.device)
        return out, kl, fa

    def _clear_taps(self):
        if self.tap_s is not None:
            self.tap_s.clear()
            self.tap_t.clear()

    def __call__(self, images, labels=None):
        total = self.compute(images)
        self.apply()
        return total

    def apply(self):
        """Gradient exchange over the data-parallel group, then the optimiser update."""
        if self.exchange:
            self.grads.all_reduce_mean(self.group)
        self.opt.step()
        # The weights just changed: the modules' cached fake-quantised copies are stale whether or not the optimiser's
        # in-place update bumped the parameters' version counters (torch's fused multi-tensor SGD does not, and a
        # CUDA-graph capture that found the cache "fresh" would leave the re-quantisation out of the graph)
        bank = _weight_bank(self.student)
        if bank is not None:
            bank.invalidate()

    def compute(self, images):
        """Both forwards and the backward of one iteration; leaves the gradients in the flat buffer."""
        self._clear_taps()
        images = images.detach().requires_grad_(True)
        t_out = self.teacher(images)
        _, kl, fa = self._losses(images, t_out)
        loss = kl + fa
        total = loss
        parts = [loss]
        if self.perturb:
            # only the images' gradient is asked for: the fused backward kernels skip their parameter-gradient
            # reductions in this sweep (the engine would discard them anyway)
            from .fusion import input_gradient_only
            with input_gradient_only():
                sign = torch.sgn(torch.autograd.grad(loss, images, retain_graph=True)[0])
            self._clear_taps()
            with torch.no_grad():
                # images + eps * sign (trainer_direct.py:510) in one launch: eps * sign is exact (sign is -1, 0, 1 or
                # NaN), so the scaled add rounds once like the two-step form and gives the same bits
                images_p = torch.add(images, sign, alpha=self.eps)
                t_out_p = self.teacher(images_p.detach())
            _, kl_p, fa_p = self._losses(images_p.detach(), t_out_p.detach())
            loss_p = kl_p + fa_p
            total = loss + loss_p
            parts.append(loss_p)
        if not self.grads.attached():
            raise RuntimeError("QATStep: parameter gradients no longer alias the flat buffer (the model was moved or "
                               "re-formatted after the step was built); create the step after model.to(...)")
        self.grads.zero()
        if self.prune_backward:
            # same pruned sweep as ``total.backward(inputs=params)``, but the gradients come back as fresh tensors and
            # land in the flat buffer with ONE multi-tensor copy instead of one ``grad += g`` launch per parameter
            # The clean and the perturbed forward are two graphs that only meet at the parameters (and at the cached
            # fake-quantised weights): one sweep over their sum makes the engine add the two gradients of every
            # parameter with its own launch (62 adds + 62 copies, 0.4 ms of the 224x224 step); a sweep per graph and two
            # multi-tensor launches give the same sums (g1 + g2, fp32 addition commutes) without them.
            for i, part in enumerate(parts):
                # (on a GPU: the ~18 BatchNorm parameter-gradient reductions of a sweep fold their per-CTA partials with
                # ONE launch when the sweep is over instead of a small launch behind each, ops.deferred_folds)
                with _deferred_folds(self.grads.flat):
                    got = torch.autograd.grad(part, self.grads.params, allow_unused=True)
                _multi_tensor_into([p.grad for p in self.grads.params], got, add=i > 0)
        else:
            total.backward()
        _drop_stem_cache()
        return total.detach()


# ------------------------------------------------------------------------------ generator warm-up phase
class GeneratorStep:
    """One iteration of the warm-up epochs 0-3 (trainer_direct.py:458-488): the phase in which the activation ranges
    are calibrated and the BN-statistics loss (trainer flavour) drives the generator.

        z, labels -> images = G(z, labels) -> teacher(images) with the BN-input hooks
        -> loss_G = CE(teacher logits, labels) + 0.1 * BNS  -> Adam step on G          (:459-487, backward_G :342-348)
        -> student(images.detach())   # no loss: every QuantAct tracks its range on 16 generated images (:488)

    ``stat`` is ``bns.BNStatLoss(teacher)`` on the GPU or the oracle's ``StatTap`` on CPU.  The generator itself is
    the reference's (main_direct.py:52-127; out of scope here, any ``G(z, labels)`` module works).  ``z`` and the labels
    are drawn on the host and moved, as in the reference, so equal seeds give equal batches on every device.
    """

    def __init__(self, generator, teacher, student, stat, latent_dim: int, n_classes: int, batch: int = 16,
                 lr: float = 1e-3, betas=(0.5, 0.999), bns_weight: float = 0.1):
        self.generator, self.teacher, self.student, self.stat = generator, teacher, student, stat
        self.latent_dim, self.n_classes, self.batch, self.bns_weight = latent_dim, n_classes, batch, bns_weight
        self.opt = torch.optim.Adam(generator.parameters(), lr=lr, betas=betas)         # trainer_direct.py:87-88
        for p in teacher.parameters():
            p.requires_grad_(False)
        student.eval()            # :411-413
        teacher.eval()
        generator.train()

    def __call__(self):
        dev = next(self.generator.parameters()).device
        z = torch.randn(self.batch, self.latent_dim).to(dev)                             # :459
        labels = torch.randint(0, self.n_classes, (self.batch,)).to(dev)                 # :460
        images = self.generator(z.contiguous(), labels.contiguous())
        self.stat.clear()
        logits = self.teacher(images)
        one_hot = F.cross_entropy(logits, labels)                                        # :471
        bns = self.stat.loss("trainer")                                                  # :473-484
        loss_g = one_hot + self.bns_weight * bns                                         # :486
        self.opt.zero_grad()
        loss_g.backward()
        self.opt.step()
        with torch.no_grad():
            self.student(images.detach())                                                # :488 (range tracking only)
        return loss_g.detach(), one_hot.detach(), bns.detach()


# ------------------------------------------------------------------------------ BN-statistics distillation
def hard_sample_loss(logits, labels, beta: float, gamma: float):
    """Focal cross-entropy of the image-distillation loop (data_generate/distill_data.py:236-249).

    ``beta * mean((1 - p_label)^gamma * CE)`` with p clamped to 1 - 1e-7; plain ``beta * mean(CE)`` for gamma = 0.
    """
    ce = F.cross_entropy(logits, labels, reduction="none")
    if gamma == 0:
        return beta * ce.mean()
    p = F.softmax(logits, dim=1).gather(1, labels.unsqueeze(1)).squeeze(1).clamp(max=1.0 - 1e-7)
    return beta * ((1 - p).pow(gamma) * ce).mean()


class PlateauOnDevice:
    """``ReduceLROnPlateau(optimizer, min_lr=1e-4, patience=50)`` of the distillation loop (distill_data.py:185-188,
    stepped at :275 with ``total_loss.item()``) without the host round trip: best value, bad-iteration counter and
    the learning rate live in device tensors and are updated by a handful of element-wise ops, so the iteration
    stays free of host synchronisation and can be replayed as a CUDA graph.  Same rule as torch's scheduler
    (mode "min", relative threshold 1e-4, factor 0.1, no cooldown, eps 1e-8), evaluated in float64 like there:
    identical decisions for identical losses (tests/test_step_cpu.py)."""

    def __init__(self, lr: torch.Tensor, factor=0.1, patience=50, threshold=1e-4, min_lr=1e-4, eps=1e-8):
        self.lr = lr                                               # the optimiser's own lr tensor, updated in place
        dev = lr.device
        self.best = torch.full((), float("inf"), dtype=torch.float64, device=dev)
        self.num_bad = torch.zeros((), dtype=torch.int64, device=dev)
        self.factor, self.patience, self.threshold, self.min_lr, self.eps = factor, patience, threshold, min_lr, eps

    def step(self, loss: torch.Tensor):
        cur = loss.detach().double().reshape(())
        better = cur < self.best * (1.0 - self.threshold)
        self.best.copy_(torch.where(better, cur, self.best))
        bad = torch.where(better, torch.zeros_like(self.num_bad), self.num_bad + 1)
        reduce = bad > self.patience
        old = self.lr.detach().double().reshape(())
        new = torch.clamp(old * self.factor, min=self.min_lr)
        apply = reduce & ((old - new) > self.eps)
        self.lr.copy_(torch.where(apply, new, old).to(self.lr.dtype).reshape(self.lr.shape))
        self.num_bad.copy_(torch.where(reduce, torch.zeros_like(bad), bad))


class DistillStep:
    """One Adam iteration on a batch of synthetic images against BN statistics (distill_data.py:229-275).

    ``stat`` is either this package's ``bns.BNStatLoss`` (GPU) or the oracle's ``StatTap`` (CPU): both expose
    ``clear()`` and ``loss(flavour)``.  The teacher may be wrapped by ``quantize_model`` first (BASELINE config 5); the
    gradient then reaches the images through cuDNN dgrad and the identity STE of every QuantAct.

    ``plateau``: the loop's ``ReduceLROnPlateau`` (:185-188, :275).  Eagerly it is torch's own scheduler fed
    ``total.item()`` as in the reference; with ``capturable`` it is ``PlateauOnDevice`` (no host sync).

    ``augment``: the per-image ``RHF(RRC(gaussian_data[j]))`` the loop puts in front of the teacher on every other
    iteration of a 224x224 batch (:197-227; RandomResizedCrop(size, scale=(augMargin, 1.0)) + RandomHorizontalFlip,
    differentiable).  A callable ``(x, boxes, flips) -> tensor``: ``augment.batch_augmenter()`` runs the whole batch
    as one kernel each way on the GPU, the CPU arm passes the oracle's torch version.  The coin (``random.random() <
    augment_p``) and the per-image draws consume Python's and torch's generators exactly like the loop, so equal
    seeds give equal crops.  Host-side randomness cannot be replayed from a CUDA graph: with ``augment`` the
    iteration runs eagerly.  ``None`` (the 28 / 32-pixel branch of the loop, and the default) applies nothing.
    """

    def __init__(self, teacher, stat, images, labels, lr=0.5, beta=0.1, gamma=0.5, capturable=False, plateau=True,
                 augment=None, augment_p=0.5, aug_margin=0.4, fused_optimizer=True):
        if augment is not None and capturable:
            raise ValueError("DistillStep: the augmented iteration draws on the host every step and cannot be captured; "
                             "use capturable=False")
        self.teacher, self.stat, self.labels, self.beta, self.gamma = teacher, stat, labels, beta, gamma
        self.augment, self.augment_p, self.aug_margin = augment, augment_p, aug_margin
        self.images = images.detach().clone().requires_grad_(True)
        # ``capturable``: keep Adam's step counter (and the learning rate) on the device so the iteration can be
        # replayed as a CUDA graph
        lr_arg = torch.tensor(float(lr), dtype=torch.float32, device=self.images.device) if capturable else lr
        # torch's single-kernel ("fused") Adam where the images live on a GPU: the update of the 154 MB batch is one
        # pass over parameter, gradient and both moments (1.1 GB) instead of the ~20 multi-tensor launches of the
        # default implementation (0.7 ms of a 9 ms iteration); same update rule, stock PyTorch either way
        extra = {"fused": True} if self.images.is_cuda and fused_optimizer else {}
        self.opt = torch.optim.Adam([self.images], lr=lr_arg, capturable=capturable, **extra)      # distill_data.py:183
        self.scheduler = None
        if plateau and capturable:
            self.scheduler = PlateauOnDevice(self.opt.param_groups[0]["lr"])
        elif plateau:
            self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.opt, min_lr=1e-4, patience=50)   # :185-188
        self._on_device = plateau and capturable
        for p in teacher.parameters():
            p.requires_grad_(False)
        teacher.eval()

    def _augmented(self, x):
        """:197-227 -- every other iteration (on average) the teacher sees a random crop / mirror of each image."""
        import random

        from .augment import random_resized_crop_params
        if random.random() < self.augment_p:
            n, _, h, w = x.shape
            boxes, flips = random_resized_crop_params(n, h, w, scale=(self.aug_margin, 1.0))
            return self.augment(x, boxes, flips)
        return x

    def __call__(self):
        self.stat.clear()
        # a fresh autograd leaf over the optimised tensor every iteration: same gradient as zero_grad() +
        # backward() on the parameter itself (:270-271), but no AccumulateGrad node tied to the stream the
        # parameter was created on, so the iteration can be captured as a CUDA graph
        x = self.images.detach().requires_grad_(True)
        # (on a GPU the ~20 statistics reductions of the forward fold their per-CTA partials with ONE launch when the
        # forward is over, before the loss reads the sums: ops.deferred_folds)
        with _deferred_folds(self.images):
            out = self.teacher(self._augmented(x) if self.augment is not None else x)
        target = hard_sample_loss(out, self.labels, self.beta, self.gamma)
        total = self.stat.loss("distill") + target                   # mean/L + var/L + target, :259-265
        self.images.grad = torch.autograd.grad(total, [x])[0]
        torch.nn.utils.clip_grad_norm_([self.images], max_norm=1.0)  # :273
        self.opt.step()
        if self.scheduler is not None:                               # :275
            self.scheduler.step(total if self._on_device else total.item())
        _drop_stem_cache()
        return total.detach()


# ------------------------------------------------------------------------------ CUDA-graph replay
class GraphedStep:
    """One whole iteration (forward, both backwards, optimiser) captured once and replayed as a CUDA graph.

    For the launch-bound configs (ResNet-20 on 32x32, ResNet-18 on 28x28) an iteration is ~2000 kernel launches
    of a few microseconds each and the host cannot enqueue them as fast as the GPU retires them; replaying the
    captured graph removes the host from the loop.  Every kernel of the library is stream-ordered and free of
    host synchronisation, so the step captures as is: the range buffers are read on the device, the per-step
    weight re-quantisation (``WeightBank``) is part of the graph.

    ``step``: a ``QATStep`` / ``DistillStep``-like callable taking one batch (or none) and returning a scalar tensor.
    """

    def __init__(self, step, example=None, warmup=3, capture_update=None):
        from .quantization_utils.quant_modules import WeightBank
        self._bank = WeightBank
        self.step = step
        self.static_in = None if example is None else example.clone()
        args = () if self.static_in is None else (self.static_in,)
        # With more than one rank the gradient all-reduce stays OUTSIDE the graph: forward and backward are
        # replayed, the NCCL collective and the (few-launch, foreach) optimiser update run eagerly behind it.
        split = hasattr(step, "compute") and hasattr(step, "apply")
        if capture_update is None:
            import torch.distributed as dist
            capture_update = not (split and dist.is_available() and dist.is_initialized()
                                  and dist.get_world_size(getattr(step, "group", None)) > 1)
        self._eager_tail = step.apply if (split and not capture_update) else None
        body = step.compute if self._eager_tail is not None else step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):              # eager warm-up off the capture stream (cuDNN autotune, workspaces)
            for _ in range(warmup):
                step(*args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._body, self._args = body, args
        # what the captured launches depend on besides tensors: the mode flags the modules branch on in Python
        self._nets = [v for v in vars(step).values() if isinstance(v, nn.Module)]
        self._flagged = [m for net in self._nets for m in net.modules() if hasattr(m, "running_stat")]
        self._capture()
        self._modes_at_capture = self._modes()

    def _modes(self):
        return (tuple(net.training for net in self._nets),
                tuple((m.running_stat, getattr(m, "full_precision_flag", False)) for m in self._flagged))

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._body(*self._args)
        self._bank.invalidate()
        self._hyper_at_capture = self._hyper()

    def _hyper(self):
        """Python-number hyper-parameters of an optimiser whose update is INSIDE the graph.  They are baked into the
        captured launches (``alpha=-lr`` of the foreach update), so ``Trainer.update_lr`` (trainer_direct.py:122-133,
        ``param_group['lr'] = ...`` once per epoch) would otherwise be ignored by every replay.  Tensor-valued ones
        (the capturable Adam of the distillation step) live on the device and need no tracking."""
        opt = getattr(self.step, "opt", None)
        if opt is None or self._eager_tail is not None:
            return None
        return [tuple((k, v) for k, v in sorted(g.items()) if k != "params" and isinstance(v, (bool, int, float)))
                for g in opt.param_groups]

    def __call__(self, batch=None, non_blocking=True):
        if self._hyper_at_capture is not None and self._hyper() != self._hyper_at_capture:
            self._capture()                                       # a learning-rate milestone: capture once more
        if self._modes() != self._modes_at_capture:
            # freeze_model / unfreeze_model / train() / eval() change which kernels a forward launches; a replay would
            # keep running the old ones (e.g. frozen ranges while the caller believes it is calibrating)
            raise RuntimeError("GraphedStep: the models' mode flags (train/eval, QuantAct.running_stat, "
                               "full_precision_flag) changed since the graph was captured; build a new GraphedStep")
        if batch is not None and self.static_in is not None:      # steps without an input (distillation) ignore it
            self.static_in.copy_(batch, non_blocking=non_blocking)
        self.graph.replay()
        if self._eager_tail is not None:
            self._eager_tail()
        # the graph re-quantised and then updated the weights behind Python's back: an eager forward after
        # this must not trust the modules' cached quantised weights
        self._bank.invalidate()
        return self.static_out
