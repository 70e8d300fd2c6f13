#!/usr/bin/env python
"""Experiment: what the ImageNet stem convolution (3 -> 64, 7x7, stride 2) costs in cuDNN with the input padded
to 4 / 8 channels (channels_last).  The convolution itself stays on cuDNN (BASELINE.json north_star); this only
measures whether the caller should hand it 16-byte pixels.

    python tools/exp_stem_conv.py
"""
import torch
import torch.nn.functional as F


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x3 = torch.randn(256, 3, 224, 224, generator=g).to(dev)
    w3 = (torch.randn(64, 3, 7, 7, generator=g) * 0.05).to(dev)
    ref = None
    for fmt_name, fmt in (("channels_last", torch.channels_last), ("nchw", torch.contiguous_format)):
        for cpad in (3, 4, 8):
            x = F.pad(x3, (0, 0, 0, 0, 0, cpad - 3)).contiguous(memory_format=fmt).requires_grad_(True)
            w = F.pad(w3, (0, 0, 0, 0, 0, cpad - 3)).contiguous(memory_format=fmt).requires_grad_(True)
            y = F.conv2d(x, w, None, 2, 3)
            if ref is None:
                ref = y.detach()
            err = (y.detach() - ref).abs().max().item()
            go = torch.randn_like(y)
            t_f = timeit(lambda: F.conv2d(x, w, None, 2, 3))
            t_d = timeit(lambda: torch.autograd.grad(y, x, go, retain_graph=True))
            t_w = timeit(lambda: torch.autograd.grad(y, w, go, retain_graph=True))
            print(f"{fmt_name:14s} C_in={cpad}: fprop {t_f:.3f} ms  dgrad {t_d:.3f} ms  wgrad {t_w:.3f} ms  "
                  f"max|y - y_ref| {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
