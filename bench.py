#!/usr/bin/env python
"""bench.py -- QAT images/sec of the OOD-DFQ quantisation path on B200 (see DESIGN.md section "Measurement").

    python bench.py --gpus 1 --steps 8 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --steps 2 --warmup 1         # the reference algorithm on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W    # weak scaling, one rank per GPU

A "step" is one steady-state data-free QAT iteration (ood_dfq_b200/step.py, reference
trainer_direct.py:490-518) of a W4A4 ResNet-18 on a synthetic 256 x 3 x 224 x 224 batch per GPU
(BASELINE.json configs[3], the configuration the metric is quoted on): teacher forward, student
forward, KD + feature-alignment loss, sign perturbation of the images, second teacher/student
forward, backward, SGD.  Activation ranges are calibrated for 3 steps first and then frozen,
as in the reference's epochs 0-3 / >= 4.

One JSON line on stdout (rank 0).  value = images/s with the batches already in HBM; e2e = the
same loop fed from pinned host memory with the loss read back every step.
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
# dram__bytes_read.sum + dram__bytes_write.sum of the largest launch of each kernel family (ncu --set full)
# (bn_* / fq_flat: the [256,64,112,112] launch; res_tail_*: [256,64,56,56]; bn_pool_fwd: [256,64,112,112] in);
# sources: profiles/r1_bn_nhwc_kernels.txt, r1_fq_flat_after_lut.txt, r1_res_tail_kernels.txt, r1_bn_pool_kernels.txt
NCU_TRAFFIC = {"bn_*_bwdx_kernel": 2431773184, "bn_*_fwd_kernel<relu,quant>": 1598203000, "bn_*_fwd_kernel": 1598398000,
               "fq_flat_kernel": 1588173312, "res_tail_bwd_kernel": 985847296, "res_tail_fwd_kernel": 580081408,
               "bn_pool_fwd_kernel": 1260052224}
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


def make_step(workload, teacher, student, namespace, group=None):
    from ood_dfq_b200 import nets, step
    lr = 1e-5 if WORKLOADS[workload][0] == "resnet20_cifar" else 1e-6      # config/*.hocon lr_S
    return step.QATStep(student, teacher, lr=lr, momentum=0.9, weight_decay=1e-4, temperature=20.0, alpha=20.0,
                        lam=1000.0, eps=0.01, unit_types=(nets.ResUnit,), group=group)


def calibrate(student, batches, namespace):
    """Three range-tracking forwards (epochs 0-3 of the reference: trainer_direct.py:488), then freeze."""
    from ood_dfq_b200 import surgery
    surgery.unfreeze_model(student, namespace)
    with torch.no_grad():
        for b in batches:
            student(b)
    surgery.freeze_model(student, namespace)


# ----------------------------------------------------------------------------- CPU arm
def run_cpu(workload, steps, warmup, sample_batch, augment=False):
    """The reference algorithm (oracle port: same ATen op sequence as the reference) on host cores."""
    from oracle import fq_torch
    torch.set_num_threads(os.cpu_count() or 1)
    _, _, shape, _, _, _ = WORKLOADS[workload]
    teacher, student = build_pair(workload, fq_torch, "cpu")
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn((sample_batch,) + shape, generator=g) for _ in range(2)]
    calibrate(student, batches[:1] * 3, fq_torch)
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
        qat = make_step(workload, teacher, student, fq_torch)
    for i in range(warmup):
        qat(batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        qat(batches[i % 2])
    dt = time.perf_counter() - t0
    return {"value": sample_batch * steps / dt, "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_step": 1e3 * dt / steps,
            "sample": f"{steps} QAT steps of {sample_batch} images ({workload}) after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU eager, oracle/fq_torch.py modules"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = run_cpu(args.workload, args.steps, args.warmup, args.cpu_batch, augment=args.distill_augment)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": res["value"], "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][5], "name": args.workload,
                   "sample_batch_per_step": args.cpu_batch,
                   "note": "reference algorithm (torch CPU eager, same op sequence) on the host cores; "
                           "each step is a bounded sample of the per-GPU batch"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def main_ours(args):
    from ood_dfq_b200 import _native, dist as ddist, ops
    from ood_dfq_b200.quantization_utils import quant_modules as qm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _native.load()
    torch.backends.cudnn.benchmark = True          # main_direct.py:351

    _, _, shape, bits, default_batch, cfg = WORKLOADS[args.workload]
    batch = args.batch or default_batch
    teacher, student = build_pair(args.workload, qm, dev)
    channels_last = not args.nchw
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
    calibrate(student, resident, qm)
    if not args.no_fuse:
        # eval-mode BN (+ ReLU + frozen QuantAct) as one kernel forward / one backward
        from ood_dfq_b200 import fusion
        fusion.fuse_eval_bn(student, resident[0][:2])
        fusion.fuse_eval_bn(teacher, resident[0][:2])
        if not args.no_tail_fuse:
            # ... and the tail of every residual unit (BN + add + ReLU + QuantAct + feature tap) as one kernel
            fusion.fuse_residual_tails(student, resident[0][:2])
            fusion.fuse_residual_tails(teacher, resident[0][:2])
        if not args.no_s2d:
            # ... and hand the 3-channel stride-2 stem convolution a space-to-depth image (cuDNN has no good
            # kernel for the 3-channel form); no-op for the networks without such a stem
            fusion.space_to_depth_stem(student, resident[0][:2])
            fusion.space_to_depth_stem(teacher, resident[0][:2])
    kind = KINDS.get(args.workload, "qat")
    if kind == "distill":
        # the "student" IS the quantised teacher here; the optimised variable is the image batch itself
        from ood_dfq_b200 import bns, step as step_mod
        del teacher
        labels = torch.randint(0, WORKLOADS[args.workload][1], (batch,), generator=g).to(dev)
        # --distill-sync (second reading): ONE global batch of world x 256 images whose BN
        # statistics are all-reduced every forward (packed fp64 partial sums over NCCL), instead of an independent
        # 256-image problem per rank.  The collective sits inside the iteration, so that mode runs eagerly.
        sync = bool(args.distill_sync and world > 1)
        if sync:
            args.graph = "off"
        aug = None
        if args.distill_augment:
            # --distill-augment: the loop's 224-pixel branch (distill_data.py:205-227) -- every other iteration each
            # image passes RandomResizedCrop(scale=(0.4, 1)) + flip before the teacher, one kernel each way for the
            # whole batch (csrc/augment.cu).  The draws happen on the host every iteration: eager launches.
            from ood_dfq_b200 import augment as augment_mod
            aug = augment_mod.batch_augmenter(channels_last)
            args.graph = "off"
        dstep = step_mod.DistillStep(student, bns.BNStatLoss(student, sync=sync), resident[0] / 5, labels,   # distill_data.py:181
                                     capturable=args.graph in ("on", "auto"), augment=aug)

        def qat(_batch=None):
            return dstep()
    else:
        qat = make_step(args.workload, teacher, student, qm)      # after .to(): gradients alias one flat buffer
    # Every workload is replayed as one CUDA graph: the small-image configs are launch-bound outright, and even
    # the 224x224 step loses ~9 % to host gaps in its 7x7 / 14x14 stages (profiles/r1_bench_n1_graph.json)
    use_graph = args.graph in ("on", "auto")
    if world > 1:
        ddist.reduce_minmax(student)
        for m in student.modules():               # ranges stay frozen from here on
            if isinstance(m, qm.QuantAct):
                m.fix()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: `value` and the fake-quant roofline ------------------------------
    for i in range(args.warmup):
        qat(resident[i % pool])
    eager_prof = None
    if use_graph:
        # launch-bound workload: the per-kernel event timing needs eager launches, so take it from a few eager
        # steps first, then capture the whole iteration as a CUDA graph and time the replays
        from ood_dfq_b200 import step as step_mod
        ops.PROFILE = []
        torch.cuda.synchronize()
        for i in range(min(args.steps, 4)):
            qat(resident[i % pool])
        torch.cuda.synchronize()
        eager_prof, ops.PROFILE = (ops.PROFILE, min(args.steps, 4)), None
        _native.reset_launch_count()
        qat(resident[0])
        launches_per_step = _native.launch_count()
        # --graph-collective (experiment, N > 1): capture the NCCL gradient all-reduce and the optimiser update inside the
        # graph as well (default: replay forward + backward, run the exchange and the update eagerly behind it)
        cap = True if (args.graph_collective and world > 1) else None
        qat = step_mod.GraphedStep(dstep, None) if kind == "distill" else \
            step_mod.GraphedStep(qat, resident[0], capture_update=cap)
    else:
        ops.PROFILE = []                            # event pairs around every streaming launch
    _native.reset_launch_count()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        qat(resident[i % pool])
    ev1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = _native.launch_count()
    ms = ev0.elapsed_time(ev1)
    prof, ops.PROFILE = ops.PROFILE, None
    prof_steps = args.steps
    if use_graph:
        launches = launches_per_step * args.steps    # kernel nodes of the library replayed inside the graph
        prof, prof_steps = eager_prof
    families = {}
    for name, a, b, nbytes in prof:
        f = families.setdefault(name, [0, 0.0, 0])
        f[0] += 1
        f[1] += a.elapsed_time(b)
        f[2] += nbytes
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * batch * args.steps / (ms / 1e3)

    # ---- end-to-end arm: pinned host batches in, loss out, every step --------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if kind == "distill":
        # an iteration has no host input (the images live on the device for the 1000 iterations of a batch,
        # distill_data.py:195); the reference reads the three loss terms back every iteration (:266-268)
        for i in range(args.warmup):
            qat().item()
        barrier()
        e0.record()
        for i in range(args.steps):
            _ = qat().item()
        e1.record()
        barrier()
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
        for i in range(args.warmup):
            qat(next(stream)[0]).item()
        barrier()
        e0.record()
        for i in range(args.steps):
            _ = qat(next(stream)[0]).item()
        e1.record()
        barrier()
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
        for i in range(args.warmup):
            _, pending = run(pending, i + 1 if i + 1 < args.warmup else None)
        barrier()
        e0.record()
        t_wall = time.perf_counter()
        pending = fetch(args.warmup)               # all K host->device copies happen inside the timed region
        for i in range(args.steps):
            j = args.warmup + i
            _, pending = run(pending, j + 1 if i + 1 < args.steps else None)
            if args.verbose and rank == 0:
                print(f"[e2e] step {i}: {1e3 * (time.perf_counter() - t_wall):.1f} ms since start", file=sys.stderr)
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * batch * args.steps / (e2e_ms / 1e3)
    h2d = 0 if kind == "distill" else batch * shape[0] * shape[1] * shape[2] * 4
    if kind != "distill" and args.e2e_input == "device_shards":
        h2d = batch * (8 + 16 + 1)                 # index, box, flip per image

    if rank == 0:
        peak, peak_src = peaks()
        table = {name: {"launches": n, "ms_per_step": t / prof_steps, "gbs": nb / (t * 1e-3) / 1e9,
                        "frac": nb / (t * 1e-3) / 1e9 / peak} for name, (n, t, nb) in families.items() if t > 0}
        dominant = max(table, key=lambda k: table[k]["ms_per_step"]) if table else None
        achieved = table[dominant]["gbs"] if dominant else None
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg, "name": args.workload, "batch_per_gpu": batch, "global_batch": batch * world,
                       "bits": bits, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2: every step streams a 154 MB batch and GBs of activations",
                       "convolutions": "cuDNN (TF32 default, as the reference)",
                       "bn_relu_quant_fusion": not args.no_fuse,
                       "residual_tail_fusion": not (args.no_fuse or args.no_tail_fuse),
                       "stem_space_to_depth": not (args.no_fuse or args.no_s2d),
                       "memory_format": "channels_last" if channels_last else "NCHW (as the reference)",
                       "cuda_graph": use_graph,
                       "graph_collective": bool(use_graph and args.graph_collective and world > 1),
                       **({"distill_batch": "global (BN statistics all-reduced)" if (args.distill_sync and world > 1)
                           else "independent per rank", "distill_augment": bool(args.distill_augment)}
                          if kind == "distill" else {})},
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "input": args.e2e_input},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": dominant,
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         # ncu --set full dram read+write of the family's largest launch, e.g. res_tail backward
                         # [256,64,56,56]: 985 847 296 B vs 1 027 604 480 algorithmic (profiles/r1_res_tail_kernels.txt)
                         "traffic": NCU_TRAFFIC.get((dominant or "").split(" ")[0]),
                         "launches_timed": len(prof), "peak_source": peak_src,
                         "kernels": table,
                         "note": "achieved = algorithmic bytes / CUDA-event time of every launch of the family inside the "
                                 "timed steps (producer-warm L2, back-to-back launches; under CUDA-graph replay: of a few "
                                 "eager steps taken just before the capture); traffic = ncu dram bytes of the family's "
                                 "largest launch, see profiles/"},
        }
        if world == 1 and not args.no_cpu_baseline:
            res = run_cpu(args.workload, args.cpu_steps, 1, args.cpu_batch, augment=args.distill_augment)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default="imagenet_resnet18_w4a4", help="one of: " + ", ".join(sorted(WORKLOADS)))
    ap.add_argument("--conf", default="", help="take the workload from a reference config/*.hocon file instead")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--cpu-batch", type=int, default=32, help="images per step of the CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="keep BatchNorm / ReLU / QuantAct as separate modules")
    ap.add_argument("--no-tail-fuse", action="store_true", help="keep the residual add and the ReLU + QuantAct behind it separate")
    ap.add_argument("--no-s2d", action="store_true", help="keep the 3-channel stem convolution in its stride-2 form")
    ap.add_argument("--nchw", action="store_true", help="keep NCHW tensors (default: channels_last memory format)")
    ap.add_argument("--e2e-input", choices=["host", "device_shards"], default="host",
                    help="end-to-end arm: pinned host batches copied every step (default, the reference's data flow) or "
                         "batches assembled on the device from an HBM-resident image set (opt-in)")
    ap.add_argument("--graph-collective", action="store_true",
                    help="N > 1: capture the gradient all-reduce and the optimiser update inside the CUDA graph (experiment)")
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
    else:
        main_ours(args)


if __name__ == "__main__":
    main()
