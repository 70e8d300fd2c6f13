#!/usr/bin/env python
"""bench.py -- QAT images/sec of the OOD-DFQ quantisation path on B200 (see DESIGN.md section "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 3                # this repo's CUDA path
    python bench.py --impl reference --steps 5 --warmup 1         # the reference's own modules on the host cores
    python bench.py --impl reference-cuda --steps 5               # ... and eagerly on the GPU (the "before")
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W    # weak scaling, one rank per GPU

A "step" is one steady-state data-free QAT iteration (ood_dfq_b200/step.py, reference
trainer_direct.py:490-518) of a W4A4 ResNet-18 on a synthetic 256 x 3 x 224 x 224 batch per GPU
(BASELINE.json configs[3], the configuration the metric is quoted on): teacher forward, student
forward, KD + feature-alignment loss, sign perturbation of the images, second teacher/student
forward, backward, SGD.  Activation ranges are calibrated for 3 steps first and then frozen,
as in the reference's epochs 0-3 / >= 4.

One JSON line on stdout (rank 0).  value = images/s with the batches already in HBM; e2e = the
same loop fed from pinned host memory with the loss read back every step.  Beside the contract's
keys the line carries
  roofline            per-family event timing of the library's kernels, traffic parsed from profiles/r2_ncu_traffic.json
  other_configs       BASELINE configs 2 (every N) and 3, 5 (N = 1) measured the same way, so the driver's runs hold them
  dp_parity           N > 1: gradient exchange, synced BN-statistics loss and reduce_minmax against the single-process
                      global batch computed on rank 0 in the same run, plus a cross-rank weight checksum
  value_unfused       N = 1: ONLY the drop-in modules (NCHW, no fusion pass): the bit-exact configuration
  gpu_eager_baseline  N = 1: the reference's own quantization_utils modules (oracle/_ref) on this GPU, eager, NCHW
  cpu_baseline        N = 1: the same modules on the host cores, same 256-image batch (also: --impl reference)
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (net factory name, num_classes, image shape, bits, default per-GPU batch, config string)
    "imagenet_resnet18_w4a4": ("resnet18_imagenet", 1000, (3, 224, 224), 4, 256,
                               "imagenet.hocon ResNet-18 W4A4 QAT step, 224x224, batch 256 per GPU"),
    "cifar100_resnet20_w4a4": ("resnet20_cifar", 100, (3, 32, 32), 4, 256,
                               "cifar100_resnet20.hocon ResNet-20 W4A4 QAT step, 32x32, batch 256 per GPU"),
    "pathmnist_resnet18_w2a2": ("resnet18_small", 9, (3, 28, 28), 2, 64,
                                "pathmnist_resnet18_w2a2.hocon ResNet-18 W2A2 QAT step, 28x28, batch 64 per GPU"),
    "distill_imagenet_resnet18_w4a4": ("resnet18_imagenet", 1000, (3, 224, 224), 4, 256,
                                       "distill_data.py BN-statistics distillation iteration, quantised ResNet-18 W4A4 "
                                       "teacher, 256x3x224x224 images per GPU (independent shard per rank)"),
}
# roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum of the dominant family's largest launch, parsed at run
# time from the committed ncu capture of THIS round's build (tools/ncu_summary.py writes the file); null when the
# family has no row there -- never a pasted constant
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")


def ncu_traffic(family):
    """(bytes, source) for a kernel family name such as ``res_tail_bwd_kernel (...)``, or (None, reason)."""
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None, "profiles/r2_ncu_traffic.json missing"
    key = (family or "").split(" ")[0]
    row = table.get("kernels", {}).get(key)
    if row is None:
        return None, f"no ncu row for {key}"
    return int(row["dram_bytes"]), f"{table.get('source', 'profiles/')}: {row.get('launch', '')}"


# BASELINE.json configs[4]: BN-statistics image distillation against a quantised teacher (distill_data.py:229-275)
KINDS = {"distill_imagenet_resnet18_w4a4": "distill"}
METRIC = "QAT images/sec ResNet-18 W4A4 224x224 (data-free QAT step, fake-quant path on sm_100a kernels)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def metric_name(workload):
    if workload == "imagenet_resnet18_w4a4":
        return METRIC
    if KINDS.get(workload) == "distill":
        return "distilled images x iterations / sec, " + WORKLOADS[workload][5]
    return "QAT images/sec " + WORKLOADS[workload][5] + " (data-free QAT step, fake-quant path on sm_100a kernels)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- model assembly
def build_pair(workload, namespace, device, seed=1):
    """(teacher, student) of one workload; student = quantize_model(teacher copy) with `namespace` classes."""
    from ood_dfq_b200 import nets, surgery
    net, classes, _, bits, _, _ = WORKLOADS[workload]
    torch.manual_seed(seed)                 # the reference seeds 1 on every rank (main_direct.py:349-350)
    teacher = getattr(nets, net)(num_classes=classes) if net != "resnet18_small" else nets.resnet18_small(3, classes)
    nets.perturb_bn_stats(teacher)
    student = surgery.quantize_model(copy.deepcopy(teacher), bits, bits, namespace=namespace)
    return teacher.to(device), student.to(device)


def make_step(workload, teacher, student, namespace, group=None, fused_attention=None):
    from ood_dfq_b200 import nets, step
    lr = 1e-5 if WORKLOADS[workload][0] == "resnet20_cifar" else 1e-6      # config/*.hocon lr_S
    return step.QATStep(student, teacher, lr=lr, momentum=0.9, weight_decay=1e-4, temperature=20.0, alpha=20.0,
                        lam=1000.0, eps=0.01, unit_types=(nets.ResUnit,), group=group, fused_attention=fused_attention)


def calibrate(student, batches, namespace):
    """Three range-tracking forwards (epochs 0-3 of the reference: trainer_direct.py:488), then freeze."""
    from ood_dfq_b200 import surgery
    surgery.unfreeze_model(student, namespace)
    with torch.no_grad():
        for b in batches:
            student(b)
    surgery.freeze_model(student, namespace)


def reference_namespace():
    """(module namespace, kind) of the reference arm: the reference's OWN ``quantization_utils`` staged unmodified under
    oracle/_ref/ by ``oracle/make_ref.py`` (``kind: "reference"``), else the oracle's op-by-op restatement
    (``kind: "port"``) -- the two places bench.py may execute anything under oracle/."""
    from oracle import make_ref
    if make_ref.available():
        return make_ref.load(), "reference"
    from oracle import fq_torch
    return fq_torch, "port"


def release():
    import gc
    from ood_dfq_b200.fusion import _S2DCache
    _S2DCache.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------- CPU arm
def run_cpu(workload, steps, warmup, sample_batch, augment=False, budget_s=None):
    """The reference's CPU implementation of the path on the host cores: its own quantization_utils modules (or the
    oracle port when they are not staged) inside the same QAT step host code, torch CPU eager, all cores."""
    ns, kind = reference_namespace()
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, shape, _, default_batch, _ = WORKLOADS[workload]
    teacher, student = build_pair(workload, ns, "cpu")
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn((sample_batch,) + shape, generator=g) for _ in range(2)]
    calibrate(student, batches[:1] * 3, ns)
    if KINDS.get(workload) == "distill":
        from ood_dfq_b200 import step as step_mod
        from oracle import bns_torch
        labels = torch.randint(0, WORKLOADS[workload][1], (sample_batch,), generator=g)
        aug = None
        if augment:                                 # the loop's 224-pixel branch (distill_data.py:205-227), torch version
            from oracle import augment_torch

            def aug(t, boxes, flips):
                return augment_torch.batch(t, range(t.shape[0]), boxes, flips, t.shape[2:], channels=3)
        dstep = step_mod.DistillStep(student, bns_torch.StatTap(student), batches[0] / 5, labels, augment=aug)

        def qat(_batch=None):
            return dstep()
    else:
        qat = make_step(workload, teacher, student, ns)
    t_w = time.perf_counter()
    for i in range(warmup):
        qat(batches[i % 2])
    per_step = (time.perf_counter() - t_w) / max(warmup, 1)
    if budget_s is not None and warmup and per_step * steps > budget_s:
        steps = max(1, int(budget_s / per_step))   # bounded sample: fewer steps of the SAME batch size
    t0 = time.perf_counter()
    for i in range(steps):
        qat(batches[i % 2])
    dt = time.perf_counter() - t0
    what = ("the reference's own quantization_utils modules (oracle/_ref, unmodified)" if kind == "reference"
            else "oracle/fq_torch.py modules (op-by-op port)")
    return {"value": sample_batch * steps / dt, "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": kind, "ms_per_step": 1e3 * dt / steps, "steps": steps, "same_config": sample_batch == default_batch,
            "sample": f"{steps} QAT steps of {sample_batch} images ({workload}) after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU eager, {what}"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch or WORKLOADS[args.workload][4]
    res = run_cpu(args.workload, args.steps, max(args.warmup, 1), batch, augment=args.distill_augment,
                  budget_s=args.cpu_budget)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": res["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": res["steps"], "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][5], "name": args.workload,
                   "batch_per_step": batch, "same_config": res["same_config"],
                   "note": "the reference's CPU path (torch CPU eager) on the host cores of the GPU box, one process; "
                           "steps are capped so that the run ends within --cpu-budget seconds"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- data-parallel parity (N > 1)
def dp_parity(dev, rank, world):
    """Multi-GPU CORRECTNESS in front of the driver: the three collectives of the path against the single-process
    GLOBAL-batch result, computed on rank 0 in the same run.

    * gradient exchange (step.FlatGrads.all_reduce_mean; reference: DDP, main_direct.py:484, trainer_direct.py:350-356):
      2 SGD steps of a W4A4 ResNet-20, each rank on its shard; oracle = rank 0 walking all shards of the global batch
      with the exchange switched off, averaging the gradients itself;
    * BN partial sums (bns.BNStatLoss(sync=True)): loss and input gradient of the global batch;
    * ``reduce_minmax`` (trainer_direct.py:368-374): the mean of the per-rank range states.
    TF32 off and cuDNN deterministic inside this block, so the only differences left are summation orders."""
    from ood_dfq_b200 import bns, dist as ddist, nets, surgery
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    cud = torch.backends.cudnn
    saved = (cud.benchmark, cud.deterministic, cud.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    cud.benchmark, cud.deterministic, cud.allow_tf32 = False, True, False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {}
    try:
        per, steps = 8, 2
        torch.manual_seed(11)
        teacher = nets.perturb_bn_stats(nets.resnet20_cifar(num_classes=10))
        student = surgery.quantize_model(copy.deepcopy(teacher), 4, 4, namespace=qm)
        teacher, student = teacher.to(dev), student.to(dev)
        g = torch.Generator().manual_seed(123)
        glob = [torch.randn(world * per, 3, 32, 32, generator=g).to(dev) for _ in range(steps + 1)]
        calibrate(student, [glob[-1][:per]] * 3, qm)             # same tensor on every rank: identical frozen ranges
        twin_t, twin_s = copy.deepcopy(teacher), copy.deepcopy(student)
        dp = make_step("cifar100_resnet20_w4a4", teacher, student, qm)
        grad_rel = weight_rel = 0.0
        if rank == 0:
            solo = make_step("cifar100_resnet20_w4a4", twin_t, twin_s, qm)
            solo.exchange = False
        for it in range(steps):
            dp.compute(ddist.shard_batch(glob[it], rank, world))
            dp.grads.all_reduce_mean(dp.group)
            got = dp.grads.flat.clone()
            dp.opt.step()
            if rank == 0:
                acc = torch.zeros_like(solo.grads.flat)
                for r in range(world):
                    solo.compute(ddist.shard_batch(glob[it], r, world))
                    acc += solo.grads.flat
                solo.grads.flat.copy_(acc / world)
                grad_rel = max(grad_rel, ((got - solo.grads.flat).abs().max() / solo.grads.flat.abs().max()).item())
                solo.opt.step()
        if rank == 0:
            a = torch.cat([p.detach().reshape(-1) for p in student.parameters()])
            b = torch.cat([p.detach().reshape(-1) for p in twin_s.parameters()])
            weight_rel = ((a - b).abs().max() / b.abs().max()).item()
        out.update(grad_max_rel=grad_rel, weights_after_2_steps_max_rel=weight_rel)
        del dp

        # BN partial sums: global-batch loss and gradient
        stat = bns.BNStatLoss(teacher, sync=True)
        x = ddist.shard_batch(glob[0], rank, world).clone().requires_grad_(True)
        teacher(x)
        loss = stat.loss()
        loss.backward()
        stat.remove()
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss.detach())
        if rank == 0:
            stat_g = bns.BNStatLoss(twin_t, sync=False)
            xg = glob[0].clone().requires_grad_(True)
            twin_t(xg)
            loss_g = stat_g.loss()
            loss_g.backward()
            stat_g.remove()
            out["bns_loss_rel"] = (abs(loss.item() - loss_g.item()) / abs(loss_g.item()))
            out["bns_grad_max_rel"] = ((x.grad - xg.grad[:per]).abs().max() / xg.grad.abs().max()).item()
            out["bns_loss_equal_on_all_ranks"] = all(torch.equal(l, losses[0]) for l in losses)

        # reduce_minmax: per-rank calibration on different data, then ONE packed all-reduce
        acts = torch.nn.Sequential(qm.QuantAct(4), qm.QuantAct(4), qm.QuantAct(8)).to(dev)
        gr = torch.Generator().manual_seed(500 + rank)
        for _ in range(2):
            acts(torch.relu(torch.randn(4, 8, 6, 6, generator=gr)).to(dev))
        before = torch.cat([torch.cat([m.x_min, m.x_max]) for m in acts])
        gathered = [torch.zeros_like(before) for _ in range(world)]
        dist.all_gather(gathered, before)
        ddist.reduce_minmax(acts)
        after = torch.cat([torch.cat([m.x_min, m.x_max]) for m in acts])
        afters = [torch.zeros_like(after) for _ in range(world)]
        dist.all_gather(afters, after)
        if rank == 0:
            want = torch.stack(gathered).sum(0) / world
            rel = ((after - want).abs() / want.abs().clamp(min=1e-12)).max().item()
            out["minmax_max_rel"] = rel
            out["minmax_equal"] = bool(rel <= 2e-7) and all(torch.equal(a, afters[0]) for a in afters)
    finally:
        cud.benchmark, cud.deterministic, cud.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    release()
    return out


def weights_checksum(model, world):
    """True when every rank holds bit-identical parameters (an all-gathered 64-bit sum of the raw bit patterns)."""
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    mine = torch.stack([flat.view(torch.int32).to(torch.int64).sum(), flat.double().abs().sum().view(torch.int64)])
    if world == 1:
        return True
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    return all(torch.equal(e, every[0]) for e in every)


# ----------------------------------------------------------------------------- GPU arm
class Env:
    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def bench_workload(args, env, workload, full=True, namespace=None, fuse=True, channels_last=True, graph=True,
                   e2e=True, fused_attention=None, steps=None, warmup=None):
    """Build, calibrate, (fuse,) warm up and time one workload on this rank's GPU.  ``full``: also the per-kernel event
    profile (roofline table).  Returns a dict; every rank returns the same timing (max over ranks)."""
    from ood_dfq_b200 import _native, dist as ddist, ops
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    ns = namespace or qm
    ours = ns is qm
    world, rank, dev = env.world, env.rank, env.dev
    steps, warmup = steps or args.steps, warmup or args.warmup
    _, _, shape, bits, default_batch, cfg = WORKLOADS[workload]
    batch = (args.batch if workload == args.workload else 0) or default_batch
    teacher, student = build_pair(workload, ns, dev)
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    if channels_last:
        # cuDNN's tensor-core convolutions are NHWC inside; channels_last tensors spare it the layout transposes
        # (28 % of the NCHW step).  Every kernel of the quantisation path takes both layouts.
        student.to(memory_format=fmt)
        teacher.to(memory_format=fmt)

    # synthetic inputs: seed = rank; a small pool of distinct batches
    g = torch.Generator().manual_seed(rank)
    pool = 3
    host = [torch.randn((batch,) + shape, generator=g).contiguous(memory_format=fmt).pin_memory() for _ in range(pool)]
    resident = [h.to(dev) for h in host]
    calibrate(student, resident, ns)
    no_tail, no_s2d = args.no_tail_fuse, args.no_s2d
    if fuse:
        # eval-mode BN (+ ReLU + frozen QuantAct) as one kernel forward / one backward
        from ood_dfq_b200 import fusion
        fusion.fuse_eval_bn(student, resident[0][:2])
        fusion.fuse_eval_bn(teacher, resident[0][:2])
        if not no_tail:
            # ... and the tail of every residual unit (BN + add + ReLU + QuantAct + feature tap) as one kernel
            fusion.fuse_residual_tails(student, resident[0][:2])
            fusion.fuse_residual_tails(teacher, resident[0][:2])
        if not no_s2d:
            # ... and hand the 3-channel stride-2 stem convolution a space-to-depth image (cuDNN has no good
            # kernel for the 3-channel form); no-op for the networks without such a stem
            fusion.space_to_depth_stem(student, resident[0][:2])
            fusion.space_to_depth_stem(teacher, resident[0][:2])
        # ... and the whole-plane average pool between the last QuantAct and Quant_Linear (bit-identical to ATen's)
        fusion.fuse_global_avgpool(student)
        fusion.fuse_global_avgpool(teacher)
    kind = KINDS.get(workload, "qat")
    graph_mode = args.graph if graph else "off"
    dstep = None
    if kind == "distill":
        # the "student" IS the quantised teacher here; the optimised variable is the image batch itself
        from ood_dfq_b200 import bns, step as step_mod
        del teacher
        labels = torch.randint(0, WORKLOADS[workload][1], (batch,), generator=g).to(dev)
        # --distill-sync (second reading): ONE global batch of world x 256 images whose BN
        # statistics are all-reduced every forward (packed fp64 partial sums over NCCL), instead of an independent
        # 256-image problem per rank.  The collective sits inside the iteration, so that mode runs eagerly.
        sync = bool(args.distill_sync and world > 1)
        if sync:
            graph_mode = "off"
        aug = None
        if args.distill_augment:
            # --distill-augment: the loop's 224-pixel branch (distill_data.py:205-227) -- every other iteration each
            # image passes RandomResizedCrop(scale=(0.4, 1)) + flip before the teacher, one kernel each way for the
            # whole batch (csrc/augment.cu).  The draws happen on the host every iteration: eager launches.
            from ood_dfq_b200 import augment as augment_mod
            aug = augment_mod.batch_augmenter(channels_last)
            graph_mode = "off"
        dstep = step_mod.DistillStep(student, bns.BNStatLoss(student, sync=sync), resident[0] / 5, labels,   # distill_data.py:181
                                     capturable=graph_mode in ("on", "auto"), augment=aug)

        def qat(_batch=None):
            return dstep()
    else:
        qat = make_step(workload, teacher, student, ns, fused_attention=fused_attention)   # after .to(): gradients alias one flat buffer
    # Every workload is replayed as one CUDA graph: the small-image configs are launch-bound outright, and even
    # the 224x224 step loses ~9 % to host gaps in its 7x7 / 14x14 stages
    use_graph = graph_mode in ("on", "auto")
    if world > 1 and ours:
        ddist.reduce_minmax(student)
        for m in student.modules():               # ranges stay frozen from here on
            if isinstance(m, qm.QuantAct):
                m.fix()

    # ---- device-resident arm: `value` and the fake-quant roofline ------------------------------
    for i in range(warmup):
        qat(resident[i % pool])
    eager_prof, launches_per_step = None, 0
    graph_collective = False
    if use_graph:
        # launch-bound workload: the per-kernel event timing needs eager launches, so take it from a few eager
        # steps first, then capture the whole iteration as a CUDA graph and time the replays
        from ood_dfq_b200 import step as step_mod
        if full:
            ops.PROFILE = []
            torch.cuda.synchronize()
            for i in range(min(steps, 4)):
                qat(resident[i % pool])
            torch.cuda.synchronize()
            eager_prof, ops.PROFILE = (ops.PROFILE, min(steps, 4)), None
        _native.reset_launch_count()
        qat(resident[0])
        launches_per_step = _native.launch_count()
        # N > 1: replay forward + backward, run the NCCL gradient all-reduce and the optimiser update eagerly behind the
        # graph.  --graph-collective captures them inside the graph as well (NCCL collectives are capturable); measured
        # at N = 2 it is the slower of the two (36.56 vs 36.32 ms at 224x224, 8.47 vs 8.37 ms at 32x32,
        # profiles/r2_n2_graph_collective.txt), so it stays an option
        graph_collective = bool(world > 1 and kind != "distill" and args.graph_collective)
        cap = True if graph_collective else None
        qat = step_mod.GraphedStep(dstep, None) if kind == "distill" else \
            step_mod.GraphedStep(qat, resident[0], capture_update=cap)
    elif full:
        ops.PROFILE = []                            # event pairs around every streaming launch
    _native.reset_launch_count()
    clocks = ClockSampler(env.local) if full else None
    env.barrier()
    if rank == 0 and clocks:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        qat(resident[i % pool])
    ev1.record()
    env.barrier()
    clk = clocks.stop() if (rank == 0 and clocks) else None
    launches = _native.launch_count()
    ms = ev0.elapsed_time(ev1)
    prof, ops.PROFILE = ops.PROFILE, None
    prof_steps = steps
    if use_graph:
        launches = launches_per_step * steps    # kernel nodes of the library replayed inside the graph
        prof, prof_steps = eager_prof if eager_prof is not None else ([], steps)
    families = {}
    for name, a, b, nbytes in (prof or []):
        f = families.setdefault(name, [0, 0.0, 0])
        f[0] += 1
        f[1] += a.elapsed_time(b)
        f[2] += nbytes
    ms = env.max_over_ranks(ms)
    res = {"workload": workload, "cfg": cfg, "batch": batch, "bits": bits, "kind": kind, "steps": steps, "warmup": warmup,
           "value": world * batch * steps / (ms / 1e3), "ms_per_step": ms / steps, "launches": int(launches),
           "use_graph": use_graph, "graph_collective": graph_collective, "clocks": clk, "families": families,
           "prof_steps": prof_steps, "prof_launches": len(prof or []), "shape": shape}

    # ---- end-to-end arm: pinned host batches in, loss out, every step --------------------------
    if e2e:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if kind == "distill":
            # an iteration has no host input (the images live on the device for the 1000 iterations of a batch,
            # distill_data.py:195); the reference reads the three loss terms back every iteration (:266-268)
            for i in range(warmup):
                qat().item()
            env.barrier()
            e0.record()
            for i in range(steps):
                _ = qat().item()
            e1.record()
            env.barrier()
        elif args.e2e_input == "device_shards":
            # OPT-IN variant of the end-to-end arm: the synthetic image set lives in HBM and every step's
            # batch is assembled on the device (gather + RandomResizedCrop + flip, csrc/augment.cu); what crosses PCIe
            # per step is the sample indices, crop boxes and flip bits drawn on the host (25 B per image), and the loss
            # on the way back.  The default arm below keeps the reference's data flow (host batches -> H2D).
            import numpy as np
            from ood_dfq_b200 import augment
            rng = np.random.default_rng(rank)
            m = max(4 * batch, 1024)
            ds = augment.DeviceShards(rng.standard_normal((m,) + shape, dtype=np.float32), rng.integers(0, 1000, m), batch,
                                      dev, rank=0, world=1, seed=rank, channels_last=channels_last, slots=2)

            def batches():
                epoch = 0
                while True:
                    ds.set_epoch(epoch)
                    yield from ds
                    epoch += 1
            stream = batches()
            for i in range(warmup):
                qat(next(stream)[0]).item()
            env.barrier()
            e0.record()
            for i in range(steps):
                _ = qat(next(stream)[0]).item()
            e1.record()
            env.barrier()
        else:
            # double-buffered prefetcher: two fixed device buffers, H2D on a copy stream while the previous step
            # computes (no allocation inside the loop: allocator traffic made this number jitter by 20 %).  Under
            # CUDA-graph replay the step copies the slot into the graph's static input (device to device) first.
            copy_stream = torch.cuda.Stream(dev)
            slots = [torch.empty_like(resident[0]) for _ in range(2)]
            released = [None, None]                    # event: the step that last read the slot has finished

            def fetch(i):
                k = i % 2
                with torch.cuda.stream(copy_stream):
                    if released[k] is not None:
                        copy_stream.wait_event(released[k])
                    slots[k].copy_(host[i % pool], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                return k, done

            def run(pending, nxt_index):
                k, done = pending
                torch.cuda.current_stream().wait_event(done)
                nxt = fetch(nxt_index) if nxt_index is not None else None   # crosses PCIe while this step computes
                loss = qat(slots[k])
                released[k] = torch.cuda.Event()
                released[k].record(torch.cuda.current_stream())
                return loss.item(), nxt                # device -> host read of the step's result

            pending = fetch(0)
            for i in range(warmup):
                _, pending = run(pending, i + 1 if i + 1 < warmup else None)
            env.barrier()
            e0.record()
            t_wall = time.perf_counter()
            pending = fetch(warmup)               # all K host->device copies happen inside the timed region
            for i in range(steps):
                j = warmup + i
                _, pending = run(pending, j + 1 if i + 1 < steps else None)
                if args.verbose and rank == 0:
                    print(f"[e2e] step {i}: {1e3 * (time.perf_counter() - t_wall):.1f} ms since start", file=sys.stderr)
            e1.record()
            env.barrier()
            del slots
        e2e_ms = env.max_over_ranks(e0.elapsed_time(e1))
        h2d = 0 if kind == "distill" else batch * shape[0] * shape[1] * shape[2] * 4
        if kind != "distill" and args.e2e_input == "device_shards":
            h2d = batch * (8 + 16 + 1)                 # index, box, flip per image
        res["e2e"] = {"value": world * batch * steps / (e2e_ms / 1e3), "unit": "images/s", "ms_per_step": e2e_ms / steps,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "input": args.e2e_input}
    if ours and kind != "distill":
        res["weights_identical_across_ranks"] = weights_checksum(student, world)
    del qat, student, host, resident
    release()
    return res


def short(res):
    """The part of a workload's result that goes under ``other_configs``."""
    out = {"workload": res["cfg"], "value": res["value"], "unit": "images/s", "ms_per_step": res["ms_per_step"],
           "batch_per_gpu": res["batch"], "steps": res["steps"], "cuda_graph": res["use_graph"],
           "graph_collective": res["graph_collective"], "gpu_launches": res["launches"]}
    if "e2e" in res:
        out["e2e"] = res["e2e"]
    if "weights_identical_across_ranks" in res:
        out["weights_identical_across_ranks"] = res["weights_identical_across_ranks"]
    return out


def main_ours(args):
    from ood_dfq_b200 import _native
    from ood_dfq_b200.quantization_utils import quant_modules as qm  # noqa: F401

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference)")
    env = Env()
    world, rank = env.world, env.rank
    torch.cuda.set_device(env.local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=env.dev)
    _native.load()
    torch.backends.cudnn.benchmark = True          # main_direct.py:351

    parity = dp_parity(env.dev, rank, world) if (world > 1 and not args.no_dp_parity) else None
    channels_last = not args.nchw
    res = bench_workload(args, env, args.workload, full=True, fuse=not args.no_fuse, channels_last=channels_last)
    if parity is not None:
        parity["weights_identical_across_ranks"] = res.get("weights_identical_across_ranks")

    # ---- the other BASELINE configs, driver-run inside the same line ----------------------------------------------
    others = {}
    if not args.no_other_configs and args.workload == "imagenet_resnet18_w4a4":
        # config 2 (cifar100, defined as a data-parallel run) at every N; configs 3 and 5 on one GPU
        names = ["cifar100_resnet20_w4a4"] + (["pathmnist_resnet18_w2a2", "distill_imagenet_resnet18_w4a4"] if world == 1 else [])
        for name in names:
            others[name] = short(bench_workload(args, env, name, full=False, fuse=not args.no_fuse,
                                                channels_last=channels_last, steps=max(args.steps, 20)))

    # ---- single-GPU companions of the headline: the bit-exact build and the reference's eager GPU path -----------
    unfused = eager = None
    if world == 1 and not args.no_baselines and KINDS.get(args.workload, "qat") == "qat":
        # (1) `value_unfused`: ONLY the drop-in modules -- NCHW like the reference, no BN / tail / stem fusion -- i.e.
        #     the configuration whose quantisation codes are bit-exact with the reference, graph replay
        u = bench_workload(args, env, args.workload, full=False, fuse=False, channels_last=False, e2e=False,
                           steps=min(args.steps, 8))
        unfused = {"value": u["value"], "unit": "images/s", "ms_per_step": u["ms_per_step"], "steps": u["steps"],
                   "config": "mirror modules only: NCHW, no fusion passes, CUDA-graph replay (bit-exact codes)"}
        # (2) `gpu_eager_baseline`: the reference's OWN modules (oracle/_ref) on this GPU, eager, NCHW, the same QAT
        #     step host code with the reference's torch expression for the feature-alignment maps: the honest
        #     "before" on identical hardware (SURVEY 8(d) last row)
        ns, kind = reference_namespace()
        e = bench_workload(args, env, args.workload, full=False, namespace=ns, fuse=False, channels_last=False,
                           graph=False, e2e=False, fused_attention=False, steps=min(args.steps, 5), warmup=3)
        eager = {"value": e["value"], "unit": "images/s", "ms_per_step": e["ms_per_step"], "steps": e["steps"],
                 "kind": kind, "same_config": True,
                 "config": "reference quantization_utils modules on cuda (six ATen passes per fake-quant, per-forward "
                           "weight re-quantisation), NCHW, eager launches, cuDNN convolutions as in the headline"}

    if rank == 0:
        peak, peak_src = peaks()
        table = {name: {"launches": n, "ms_per_step": t / res["prof_steps"], "gbs": nb / (t * 1e-3) / 1e9,
                        "frac": nb / (t * 1e-3) / 1e9 / peak} for name, (n, t, nb) in res["families"].items() if t > 0}
        dominant = max(table, key=lambda k: table[k]["ms_per_step"]) if table else None
        achieved = table[dominant]["gbs"] if dominant else None
        traffic, traffic_src = ncu_traffic(dominant)
        kind = res["kind"]
        line = {
            "metric": metric_name(args.workload), "value": res["value"], "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": res["cfg"], "name": args.workload, "batch_per_gpu": res["batch"],
                       "global_batch": res["batch"] * world, "bits": res["bits"], "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2: every step streams a 154 MB batch and GBs of activations",
                       "convolutions": "cuDNN (TF32 default, as the reference)",
                       "bn_relu_quant_fusion": not args.no_fuse,
                       "residual_tail_fusion": not (args.no_fuse or args.no_tail_fuse),
                       "stem_space_to_depth": not (args.no_fuse or args.no_s2d),
                       "global_avgpool_kernel": not args.no_fuse,
                       "deferred_param_grad_folds": not os.environ.get("OODFQ_NO_DEFERRED_FOLDS"),
                       "memory_format": "channels_last" if channels_last else "NCHW (as the reference)",
                       "cuda_graph": res["use_graph"], "graph_collective": res["graph_collective"],
                       **({"distill_batch": "global (BN statistics all-reduced)" if (args.distill_sync and world > 1)
                           else "independent per rank", "distill_augment": bool(args.distill_augment)}
                          if kind == "distill" else {})},
            "e2e": res["e2e"],
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "roofline": {"bound": "hbm", "kernel": dominant,
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "launches_timed": res["prof_launches"], "peak_source": peak_src,
                         "kernels": table,
                         "note": "achieved = algorithmic bytes / CUDA-event time of every launch of the family inside the "
                                 "timed steps (producer-warm L2, back-to-back launches; under CUDA-graph replay: of a few "
                                 "eager steps taken just before the capture); traffic = ncu dram bytes of the family's "
                                 "largest launch, parsed from profiles/r2_ncu_traffic.json"},
        }
        if parity is not None:
            line["dp_parity"] = parity
        if others:
            line["other_configs"] = others
        if unfused is not None:
            line["value_unfused"] = unfused
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if world == 1 and not args.no_cpu_baseline:
            cpu_batch = args.cpu_batch or res["batch"]
            c = run_cpu(args.workload, args.cpu_steps, 1, cpu_batch, augment=args.distill_augment, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main_reference_cuda(args):
    """``--impl reference-cuda``: only the GPU-eager reference leg, as its own JSON line (rank 0 of a single process)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: --impl reference-cuda needs a CUDA device")
    env = Env()
    torch.cuda.set_device(env.local)
    env.world = 1
    torch.backends.cudnn.benchmark = True
    ns, kind = reference_namespace()
    e = bench_workload(args, env, args.workload, full=False, namespace=ns, fuse=False, channels_last=False, graph=False,
                       e2e=False, fused_attention=False)
    emit({"impl": "reference-cuda", "metric": metric_name(args.workload), "value": e["value"], "unit": "images/s",
          "n_gpus": 1, "steps": e["steps"], "warmup": e["warmup"], "ms_per_step": e["ms_per_step"],
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": e["cfg"], "name": args.workload, "batch_per_gpu": e["batch"], "kind": kind,
                     "note": "reference quantization_utils modules on cuda, NCHW, eager, same QAT step host code"},
          "gpu_launches": 0})


def workload_from_conf(path):
    """Register a workload described by one of the reference's config/*.hocon files (verbatim settings:
    the file's own qw/qa/batchSize, not the W4A4 / batch-256 overrides BASELINE.json quotes the metric on)."""
    from ood_dfq_b200 import hocon
    s = hocon.QuantSettings.from_file(path)
    if s.img_size == 28:
        net = "resnet18_small"
    elif s.model_name.startswith("resnet20"):
        net = "resnet20_cifar"
    elif s.model_name == "resnet18":
        net = "resnet18_imagenet"
    else:
        raise SystemExit(f"bench.py: no carrier network for model_name={s.model_name!r} (see ood_dfq_b200/nets.py)")
    if s.qw != s.qa:
        raise SystemExit("bench.py: the workload table assumes qw == qa")
    name = "conf:" + os.path.basename(path)
    WORKLOADS[name] = (net, s.nClasses, (s.channels, s.img_size, s.img_size), s.qw, s.batchSize,
                       f"{os.path.basename(path)} {s.model_name} W{s.qw}A{s.qa} QAT step, {s.img_size}x{s.img_size}, "
                       f"batch {s.batchSize} per GPU (file values)")
    return name


_REAL_STDOUT = None


def quiet_stdout():
    """Point file descriptor 1 at stderr for the rest of the run: libraries write banners straight to it (NCCL
    prints its version line on communicator creation) and the contract is ONE JSON line on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference", "reference-cuda"], default="ours")
    ap.add_argument("--workload", default="imagenet_resnet18_w4a4", help="one of: " + ", ".join(sorted(WORKLOADS)))
    ap.add_argument("--conf", default="", help="take the workload from a reference config/*.hocon file instead")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="images per step of the CPU arm (default: the workload's per-GPU batch, i.e. the same config)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--cpu-budget", type=float, default=240.0,
                    help="CPU arm: cap the number of timed steps so that they fit this many seconds")
    ap.add_argument("--no-baselines", action="store_true", help="skip the value_unfused and gpu_eager_baseline legs (N = 1)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE configs (other_configs)")
    ap.add_argument("--no-dp-parity", action="store_true", help="N > 1: skip the data-parallel parity block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="keep BatchNorm / ReLU / QuantAct as separate modules")
    ap.add_argument("--no-tail-fuse", action="store_true", help="keep the residual add and the ReLU + QuantAct behind it separate")
    ap.add_argument("--no-s2d", action="store_true", help="keep the 3-channel stem convolution in its stride-2 form")
    ap.add_argument("--nchw", action="store_true", help="keep NCHW tensors (default: channels_last memory format)")
    ap.add_argument("--e2e-input", choices=["host", "device_shards"], default="host",
                    help="end-to-end arm: pinned host batches copied every step (default, the reference's data flow) or "
                         "batches assembled on the device from an HBM-resident image set (opt-in)")
    ap.add_argument("--graph-collective", action="store_true",
                    help="N > 1: capture the gradient all-reduce and the optimiser update inside the CUDA graph too "
                         "(default: eager behind the replay, which measured faster)")
    ap.add_argument("--distill-augment", action="store_true",
                    help="distillation workload: apply the loop's per-image RandomResizedCrop / flip on every other "
                         "iteration (distill_data.py:205-227; eager)")
    ap.add_argument("--distill-sync", action="store_true",
                    help="distillation workload at N > 1: one global batch with all-reduced BN statistics (eager)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--graph", choices=["auto", "on", "off"], default="auto",
                    help="replay the whole iteration as a CUDA graph (auto = on; off: eager launches)")
    args = ap.parse_args()
    if args.conf:
        args.workload = workload_from_conf(args.conf)
    if args.workload not in WORKLOADS:
        raise SystemExit(f"bench.py: unknown workload {args.workload!r}")
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        main_reference(args)
    elif args.impl == "reference-cuda":
        main_reference_cuda(args)
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
