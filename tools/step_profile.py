#!/usr/bin/env python
"""Per-kernel share of one QAT step (CUPTI via torch.profiler): which kernels the step spends its time in.

    python tools/step_profile.py [--workload imagenet_resnet18_w4a4] [--steps 2] [--out profiles/step_share.txt]

ncu serialises every launch and replays it, which on a 150 ms step with ~1500 launches and 30 GB of live
activations does not finish; the launch list of OUR kernels comes from ncu (profiles/*launches*.csv), the
share of the whole step from this CUPTI trace.  Numbers here are device-side kernel durations.
"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="imagenet_resnet18_w4a4")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--nchw", action="store_true")
    ap.add_argument("--no-tail-fuse", action="store_true")
    ap.add_argument("--no-s2d", action="store_true")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by-shape", action="store_true",
                    help="instead of the CUPTI table: CUDA-event pairs around every launch of the library (ops.PROFILE), "
                         "grouped by (kernel family, algorithmic bytes) = one row per tensor size")
    args = ap.parse_args()
    from ood_dfq_b200.quantization_utils import quant_modules as qm
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    _, _, shape, bits, default_batch, cfg = bench.WORKLOADS[args.workload]
    batch = args.batch or default_batch
    teacher, student = bench.build_pair(args.workload, qm, dev)
    fmt = torch.contiguous_format if args.nchw else torch.channels_last
    student.to(memory_format=fmt)
    teacher.to(memory_format=fmt)
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn((batch,) + shape, generator=g).to(dev).contiguous(memory_format=fmt) for _ in range(2)]
    bench.calibrate(student, xs + xs[:1], qm)
    if not args.no_fuse:
        from ood_dfq_b200 import fusion
        fusion.fuse_eval_bn(student, xs[0][:2])
        fusion.fuse_eval_bn(teacher, xs[0][:2])
        if not args.no_tail_fuse:
            fusion.fuse_residual_tails(student, xs[0][:2])
            fusion.fuse_residual_tails(teacher, xs[0][:2])
        if not args.no_s2d:
            fusion.space_to_depth_stem(student, xs[0][:2])
            fusion.space_to_depth_stem(teacher, xs[0][:2])
        fusion.fuse_global_avgpool(student)
        fusion.fuse_global_avgpool(teacher)
    if bench.KINDS.get(args.workload) == "distill":
        from ood_dfq_b200 import bns, step as step_mod
        labels = torch.randint(0, bench.WORKLOADS[args.workload][1], (batch,), generator=g).to(dev)
        dstep = step_mod.DistillStep(student, bns.BNStatLoss(student), xs[0] / 5, labels)

        def qat(_batch=None):
            return dstep()
    else:
        qat = bench.make_step(args.workload, teacher, student, qm)
    for i in range(3):
        qat(xs[i % 2])
    torch.cuda.synchronize()
    if args.by_shape:
        from ood_dfq_b200 import ops
        peak = bench.peaks()[0]
        ops.PROFILE = []
        for i in range(args.steps):
            qat(xs[i % 2])
        torch.cuda.synchronize()
        rec, ops.PROFILE = ops.PROFILE, None
        agg = defaultdict(lambda: [0.0, 0])
        for name, e0, e1, nbytes in rec:
            a = agg[(name.split(" (")[0], nbytes)]
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
        total = sum(a[0] for a in agg.values())
        lines = [f"# {cfg}; batch {batch}; {args.steps} eager steps; event pairs around the library's launches: {total / args.steps:.2f} ms/step",
                 f"# {'ms/step':>8s} {'launches':>9s} {'MB/launch':>10s} {'us/launch':>10s} {'GB/s':>7s} {'of peak':>8s}  kernel family"]
        for (name, nbytes), (ms, n) in sorted(agg.items(), key=lambda kv: (kv[0][0], -kv[0][1])):
            gbs = nbytes * n / (ms * 1e-3) / 1e9
            lines.append(f"  {ms / args.steps:8.3f} {n / args.steps:9.1f} {nbytes / 1e6:10.1f} {1e3 * ms / n:10.1f} {gbs:7.0f} {100 * gbs / peak:7.1f}%  {name}")
        text = "\n".join(lines)
        print(text)
        if args.out:
            with open(args.out, "w") as f:
                f.write(text + "\n")
        return
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            qat(xs[i % 2])
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
            agg[ev.name][0] += t
            agg[ev.name][1] += 1
    total = sum(v[0] for v in agg.values())
    lines = [f"# {cfg}; batch {batch}; {args.steps} steps; device kernel time {total / args.steps / 1e3:.2f} ms/step",
             f"# {'share':>6s} {'ms/step':>9s} {'launches/step':>14s}  kernel"]
    ours = 0.0
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.top]:
        mine = "oodfq::" in name
        ours += t if mine else 0.0
        lines.append(f"  {100 * t / total:6.2f} {t / args.steps / 1e3:9.3f} {n / args.steps:14.1f}  {'*' if mine else ' '} {name[:110]}")
    ours = sum(t for name, (t, n) in agg.items() if "oodfq::" in name)
    lines.append(f"# kernels of liboodfq_b200.so (*): {100 * ours / total:.2f}% of device time, {ours / args.steps / 1e3:.3f} ms/step")
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
