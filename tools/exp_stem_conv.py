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


def s2d_input(x, cpad=12):
    """[N,3,H,W] -> pad 3 -> space-to-depth by 2 -> [N,12,(H+6)/2,(W+6)/2] (channel order c,s,t), zero-padded to cpad."""
    n, c, h, w = x.shape
    xp = F.pad(x, (3, 3, 3, 3))
    hp, wp = (h + 6) // 2, (w + 6) // 2
    xs = xp.view(n, c, hp, 2, wp, 2).permute(0, 1, 3, 5, 2, 4).reshape(n, c * 4, hp, wp)
    if cpad > c * 4:
        xs = F.pad(xs, (0, 0, 0, 0, 0, cpad - c * 4))
    return xs.contiguous(memory_format=torch.channels_last)


def s2d_weight(w, cpad=12):
    """[O,3,7,7] -> pad to 8x8 -> [O,12,4,4] with w2[o,(c,s,t),p,q] = w[o,c,2p+s,2q+t]."""
    o, c, _, _ = w.shape
    w8 = F.pad(w, (0, 1, 0, 1))
    w2 = w8.view(o, c, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(o, c * 4, 4, 4)
    if cpad > c * 4:
        w2 = F.pad(w2, (0, 0, 0, 0, 0, cpad - c * 4))
    return w2.contiguous(memory_format=torch.channels_last)


def main_s2d():
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x3 = torch.randn(256, 3, 224, 224, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    w3 = (torch.randn(64, 3, 7, 7, generator=g) * 0.05).to(dev)
    ref = F.conv2d(x3, w3, None, 2, 3)
    for cpad in (12, 16):
        x = x3.clone().requires_grad_(True)
        w = w3.clone().requires_grad_(True)
        xs, ws = s2d_input(x, cpad), s2d_weight(w, cpad)
        y = F.conv2d(xs, ws, None, 1, 0)
        err = (y.detach() - ref).abs().max().item()
        go = torch.randn_like(y)
        xs_l, ws_l = xs.detach().requires_grad_(True), ws.detach().requires_grad_(True)
        y_l = F.conv2d(xs_l, ws_l, None, 1, 0)
        t_f = timeit(lambda: F.conv2d(xs_l, ws_l, None, 1, 0))
        t_d = timeit(lambda: torch.autograd.grad(y_l, xs_l, go, retain_graph=True))
        t_w = timeit(lambda: torch.autograd.grad(y_l, ws_l, go, retain_graph=True))
        t_in = timeit(lambda: s2d_input(x3, cpad))
        t_full_d = timeit(lambda: torch.autograd.grad(y, x, go, retain_graph=True))
        print(f"space-to-depth C_in={cpad}: fprop {t_f:.3f} ms  dgrad {t_d:.3f} ms  wgrad {t_w:.3f} ms  "
              f"input transform {t_in:.3f} ms  dgrad incl. inverse transform {t_full_d:.3f} ms  "
              f"shape {tuple(y.shape)}  max|y - y_ref| {err:.2e}", flush=True)


if __name__ == "__main__":
    main_s2d()
    main()
