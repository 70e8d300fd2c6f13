#!/usr/bin/env python
"""Per-kernel bandwidth microbenchmark at the BASELINE tensor shapes.

    python tools/microbench.py [--only fq,calib,minmax,stats,stats_fq,bwd,weights,copy,augment,...] [--iters 10] [--json out.json]

Timing hygiene: 3 warm-up launches, L2 flushed between timed launches (a 512 MB memset-like
write), CUDA events on the launching stream, median over --iters.  GB/s = algorithmic bytes /
time; fractions are of MEASURED_PEAKS.json hbm_gbs when present.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ood_dfq_b200 import ops  # noqa: E402

ACT_SHAPES = [(256, 64, 112, 112), (256, 64, 56, 56), (256, 128, 28, 28), (256, 256, 14, 14), (256, 512, 7, 7),
              (256, 16, 32, 32), (256, 32, 16, 16), (256, 64, 8, 8)]      # 0-4: ResNet-18 / 224; 5-7: ResNet-20 / 32 (--shapes picks)
R18_WEIGHTS = [(64, 3, 7, 7)] + [(64, 64, 3, 3)] * 4 + [(128, 64, 3, 3), (128, 128, 3, 3), (128, 64, 1, 1)] + \
              [(128, 128, 3, 3)] * 2 + [(256, 128, 3, 3), (256, 256, 3, 3), (256, 128, 1, 1)] + [(256, 256, 3, 3)] * 2 + \
              [(512, 256, 3, 3), (512, 512, 3, 3), (512, 256, 1, 1)] + [(512, 512, 3, 3)] * 2 + [(1000, 512)]


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


class Timer:
    """flush = "write": fill a 512 MB buffer (L2 left full of DIRTY lines whose write-back then competes with the
    timed kernel -- the harsher convention); "read": sum a 512 MB buffer (L2 left full of clean foreign lines)."""

    def __init__(self, iters, flush="write", warmup=3, use_cupti=False):
        self.iters, self.mode, self.warmup, self.use_cupti = iters, flush, warmup, use_cupti
        self.flush = torch.ones(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda") if flush != "none" else None

    def cupti(self, fn):
        """Device-side kernel durations from CUPTI (torch.profiler) instead of CUDA events: event timestamps tick every
        ~2 us on this GPU, too coarse for the 5-20 us launches of the small planes.  The time of a call is the sum of
        the kernels it launched (launch gaps excluded, as in a CUDA-graph replay); read-flush only."""
        from torch.profiler import ProfilerActivity, profile
        for _ in range(self.warmup):
            fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(self.iters):
                if self.flush is not None:
                    self.flush.sum()
                fn()
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                     key=lambda e: e.time_range.start)
        times, cur = [], None
        for e in evs:
            if "reduce_kernel" in e.name:                 # the flush: a new timed call starts behind it
                if cur:
                    times.append(cur)
                cur = 0.0
                continue
            if "Memset" in e.name:
                continue
            if cur is None:
                cur = 0.0
            cur += (getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)) * 1e-3
        if cur:
            times.append(cur)
        if self.flush is None:                            # no separators: everything is one segment
            times = [times[0] / self.iters] if times else [0.0]
        return statistics.median(times), min(times)

    def __call__(self, fn):
        if self.use_cupti:
            return self.cupti(fn)
        for _ in range(self.warmup):
            fn()
        times = []
        for _ in range(self.iters):
            if self.flush is not None:
                if self.mode == "write":
                    self.flush.fill_(1.0)
                else:
                    self.flush.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        return statistics.median(times), min(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="copy,fq,minmax,calib,stats,stats_fq,bwd,weights")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default="")
    ap.add_argument("--shapes", default="", help="indices into the activation shape list, e.g. 0,4")
    ap.add_argument("--flush", choices=["write", "read", "none"], default="write")
    ap.add_argument("--cupti", action="store_true", help="kernel durations from CUPTI instead of CUDA events (needs --flush read or none)")
    ap.add_argument("--warmup", type=int, default=3, help="untimed launches per kernel (0 under ncu: one launch per variant)")
    args = ap.parse_args()
    only = set(args.only.split(","))
    shapes = ACT_SHAPES[:5] if not args.shapes else [ACT_SHAPES[int(i)] for i in args.shapes.split(",")]
    pk = peak()
    if args.cupti and args.flush == "write":
        raise SystemExit("--cupti separates the timed calls by the read-flush kernel: use --flush read (or none)")
    timer = Timer(args.iters, flush=args.flush, warmup=args.warmup, use_cupti=args.cupti)
    print(f"# L2 flush between timed launches: {args.flush}; timing: {'CUPTI kernel durations' if args.cupti else 'CUDA events'}", flush=True)
    rows = []

    def report(kernel, shape, nbytes, med, best):
        gbs = nbytes / (med * 1e-3) / 1e9
        rows.append({"kernel": kernel, "shape": list(shape), "bytes": nbytes, "ms_median": med, "ms_best": best,
                     "gbs": gbs, "frac_of_peak": gbs / pk})
        print(f"{kernel:11s} {str(tuple(shape)):24s} {nbytes / 1e6:9.1f} MB  {med * 1e3:9.1f} us  "
              f"{gbs:7.0f} GB/s  {100 * gbs / pk:5.1f}% of {pk:.0f}", flush=True)

    torch.manual_seed(0)
    for shape in shapes:
        x = torch.relu(torch.randn(shape, device="cuda"))
        n = x.numel()
        lo, hi = torch.zeros(1, device="cuda"), torch.full((1,), 2.5, device="cuda")
        if "copy" in only:
            y = torch.empty_like(x)
            report("copy", shape, 8 * n, *timer(lambda: y.copy_(x)))
            del y
        if "fq" in only:
            report("fq", shape, 8 * n, *timer(lambda: ops.fake_quant(x, 4, lo, hi)))
        if "minmax" in only:
            report("minmax", shape, 4 * n, *timer(lambda: ops.minmax(x)))
        if "calib" in only:
            st = [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda"),
                  torch.full((1,), 0.9, device="cuda"), torch.ones(1, device="cuda")]

            def calib():
                st[3].fill_(1.0)
                ops.act_calib_forward(x, 4, st[0], st[1], st[2], st[3])
            report("calib", shape, 12 * n, *timer(calib))

            def calib2():
                st[3].fill_(1.0)
                ops.act_calib_forward(x, 4, st[0], st[1], st[2], st[3], onchip=False)
            report("calib_2k", shape, 12 * n, *timer(calib2))
        if "calib_stats" in only:
            # north_star (b): range update + fake-quant + per-channel sums.  On chip (<= 96 MB) x crosses HBM once:
            # 8 B/elem; above that the floor is 12 B/elem (the fraction printed is against the bytes stated here)
            st = [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda"),
                  torch.full((1,), 0.9, device="cuda"), torch.ones(1, device="cuda")]
            per = 8 if 4 * n <= (96 << 20) else 12
            for fmt, tag in ((torch.contiguous_format, "nchw"), (torch.channels_last, "nhwc")):
                xf = x.contiguous(memory_format=fmt)
                sums = torch.empty(2 * shape[1], dtype=torch.float64, device="cuda")

                def cs(onchip=True):
                    st[3].fill_(1.0)
                    ops.act_calib_stats_forward(xf, 4, st[0], st[1], st[2], st[3], sums=sums, onchip=onchip)
                report(f"cstat_{tag}", shape, per * n, *timer(cs))
                report(f"cstat2_{tag}", shape, 12 * n, *timer(lambda: cs(False)))
                del xf
        c = shape[1]
        shift = torch.zeros(c, device="cuda")
        if "stats" in only:
            report("stats", shape, 4 * n, *timer(lambda: ops.bn_stats_forward(x, shift)))
        if "stats_fq" in only:
            report("stats_fq", shape, 8 * n, *timer(lambda: ops.bn_stats_forward(x, shift, fq=(4, lo, hi))))
        if "bwd" in only:
            g = torch.randn_like(x)
            mean = torch.zeros(c, device="cuda")
            gm, gv = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
            m = float(n // c)
            report("bwd", shape, 12 * n, *timer(lambda: ops.bn_stats_backward(x, g, mean, gm, gv, m)))
            del g
        if "stats_nhwc" in only:
            xl = x.contiguous(memory_format=torch.channels_last)
            report("st_nhwc", shape, 4 * n, *timer(lambda: ops.bn_stats_forward(xl, shift)))
            report("stq_nhwc", shape, 8 * n, *timer(lambda: ops.bn_stats_forward(xl, shift, fq=(4, lo, hi))))
            gl = torch.randn_like(xl)
            mean0 = torch.zeros(c, device="cuda")
            gm0, gv0 = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
            report("bwd_nhwc", shape, 12 * n, *timer(lambda: ops.bn_stats_backward(xl, gl, mean0, gm0, gv0, float(n // c))))
            del xl, gl
        if "bn_fwd" in only or "bn_bwd" in only:
            w, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
            rm, rv = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
            for fmt, tag in ((torch.contiguous_format, "nchw"), (torch.channels_last, "nhwc")):
                xf = torch.randn(shape, device="cuda").contiguous(memory_format=fmt)
                if "bn_fwd" in only:
                    report(f"bnq_{tag}", shape, 8 * n, *timer(lambda: ops.bn_eval_forward(xf, w, b, rm, rv, 1e-5, relu=True, fq=(4, lo, hi))))
                    report(f"bn_{tag}", shape, 8 * n, *timer(lambda: ops.bn_eval_forward(xf, w, b, rm, rv, 1e-5)))
                if "bn_bwd" in only:
                    gf = torch.randn(shape, device="cuda").contiguous(memory_format=fmt)
                    report(f"bnbw_{tag}", shape, 12 * n, *timer(lambda: ops.bn_eval_backward(xf, gf, w, b, rm, rv, 1e-5, relu=True)))
                    report(f"bnbx_{tag}", shape, 8 * n, *timer(lambda: ops.bn_eval_backward(xf, gf, w, b, rm, rv, 1e-5, relu=False, want_param_grads=False)))
                    del gf
                del xf
        if "tail" in only and shape[2] <= 56:
            def bn():
                return (torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3,
                        torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5, 1e-5)
            bn1, bn2 = bn(), bn()
            x1 = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
            r = torch.relu(torch.randn(shape, device="cuda")).contiguous(memory_format=torch.channels_last)
            gy = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
            ge = torch.randn(shape[:2], device="cuda")
            report("tail_fq_e", shape, 12 * n, *timer(lambda: ops.res_tail_forward(x1, r, bn1, None, fq=(4, lo, hi), want_energy=True)))
            report("tail_fq_id", shape, 12 * n, *timer(lambda: ops.res_tail_forward(x1, r, bn1, bn2, fq=(4, lo, hi), want_energy=True)))
            report("tail_plain", shape, 12 * n, *timer(lambda: ops.res_tail_forward(x1, r, bn1, None, want_energy=True)))
            report("tailbw_ep", shape, 20 * n, *timer(lambda: ops.res_tail_backward(gy, ge, x1, r, bn1, None)))
            report("tailbw_idp", shape, 20 * n, *timer(lambda: ops.res_tail_backward(gy, ge, x1, r, bn1, bn2)))
            report("tailbw_e", shape, 20 * n, *timer(lambda: ops.res_tail_backward(gy, ge, x1, r, bn1, None, want_param_grads=False)))
            _, _, mask = ops.res_tail_forward(x1, r, bn1, None, fq=(4, lo, hi), want_energy=True, want_mask=True)
            report("tail_fq_em", shape, 12 * n + n // 4, *timer(lambda: ops.res_tail_forward(x1, r, bn1, None, fq=(4, lo, hi), want_energy=True, want_mask=True)))
            report("tailbwm_ep", shape, 16 * n + n // 4, *timer(lambda: ops.res_tail_backward(gy, ge, x1, None, bn1, None, mask=mask)))
            report("tailbwm_e", shape, 16 * n + n // 4, *timer(lambda: ops.res_tail_backward(gy, ge, x1, None, bn1, None, want_param_grads=False, mask=mask)))
            report("tailbwm_e2", shape, 20 * n + n // 4, *timer(lambda: ops.res_tail_backward(gy, ge, x1, None, bn1, None, want_param_grads=False, mask=mask, grad_y2=gy)))
            del mask
            del x1, r, gy
        if "pool" in only and shape[2] >= 56:
            w, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
            rm, rv = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
            xf = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
            out, idx, xhat = ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, fq=(4, lo, hi))
            no = out.numel()
            report("pool_fqx", shape, 4 * n + 9 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, fq=(4, lo, hi))))
            report("pool_fq", shape, 4 * n + 5 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, fq=(4, lo, hi), want_xhat=False)))
            report("pool_plain", shape, 4 * n + 5 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, want_xhat=False)))
            # the register-staged fallback (round-2 default until the TMA ring)
            report("poolr_fqx", shape, 4 * n + 9 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, fq=(4, lo, hi), register_kernel=True)))
            report("poolr_fq", shape, 4 * n + 5 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, fq=(4, lo, hi), want_xhat=False, register_kernel=True)))
            report("poolr_plain", shape, 4 * n + 5 * no, *timer(lambda: ops.bn_pool_forward(xf, w, b, rm, rv, 1e-5, want_xhat=False, register_kernel=True)))
            go = torch.randn_like(out)
            report("poolbw_p", shape, 4 * n + 9 * no, *timer(lambda: ops.bn_pool_backward(go, idx, xhat, xf.shape, w, b, rm, rv, 1e-5)))
            report("poolbw_x", shape, 4 * n + 5 * no, *timer(lambda: ops.bn_pool_backward(go, idx, None, xf.shape, w, b, rm, rv, 1e-5, want_param_grads=False)))
            del xf, out, idx, xhat, go
        del x
        torch.cuda.empty_cache()
    if "augment" in only:
        from ood_dfq_b200 import augment
        for (m, c, side, batch) in ((4096, 3, 224, 256), (20000, 3, 32, 256), (20000, 1, 28, 64)):
            images = torch.randn((m, c, side, side), device="cuda")
            gen = torch.Generator().manual_seed(0)
            index = torch.randint(0, m, (batch,), generator=gen).cuda()
            boxes, flips = augment.random_resized_crop_params_batched(batch, side, side, generator=gen)
            read = int(4 * c * (boxes[:, 2].astype("int64") * boxes[:, 3]).sum())
            boxes, flips = torch.from_numpy(boxes).cuda(), torch.from_numpy(flips).cuda()
            for cl, tag in ((True, "aug_nhwc"), (False, "aug_nchw")):
                out = torch.empty((batch, 3, side, side), device="cuda",
                                  memory_format=torch.channels_last if cl else torch.contiguous_format)
                report(tag, (batch, c, side, side), read + 4 * out.numel(),
                       *timer(lambda: ops.crop_resize_flip(images, index, boxes, flips, side, channels_last=cl, out=out)))
            del images
            torch.cuda.empty_cache()
    if "s2d" in only:
        # stem input re-layout (csrc/s2d_stem.cu): the 154 MB image batch -> 2x2 space-to-depth (12 channels), and back
        xi = torch.randn(256, 3, 224, 224, device="cuda").contiguous(memory_format=torch.channels_last)
        xs = ops.s2d_stem_forward(xi, 3)
        gs = torch.randn_like(xs)
        report("s2d_fwd", (256, 3, 224, 224), 4 * (xi.numel() + xs.numel()), *timer(lambda: ops.s2d_stem_forward(xi, 3)))
        report("s2d_bwd", (256, 3, 224, 224), 4 * (xi.numel() + xs.numel()), *timer(lambda: ops.s2d_stem_backward(gs, xi.shape, 3)))
        del xi, xs, gs
    if "weights" in only:
        ws = [torch.randn(s, device="cuda") * 0.02 for s in R18_WEIGHTS]
        outs = [torch.empty_like(w) for w in ws]
        n = sum(w.numel() for w in ws)
        ks, sym = [4] * len(ws), [False] * len(ws)
        report("weights", (len(ws), n), 8 * n, *timer(lambda: ops.weight_fq_multi(ws, ks, sym, outs=outs)))
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"peak_gbs": pk, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
