#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference, which does not exist
on the GPU box).  The committed .npz files are what travels.  Every expected
value is produced by importing ``quantization_utils`` (and the BN hook of
``data_generate/distill_data.py``) from the reference tree; nothing from this
repository's oracle or kernels is involved.

    python tools/make_golden.py            # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("OODFQ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

from quantization_utils import quant_utils as RU  # noqa: E402
from quantization_utils import quant_modules as RM  # noqa: E402
from data_generate.distill_data import DistillData  # noqa: E402

torch.set_num_threads(1)


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes, {len(arrays)} arrays")


def ref_codes(x, k, lo, hi):
    # the reference never exposes the codes; this is quant_utils.py:148-152 verbatim calls
    s, z = RU.asymmetric_linear_quantization_params(k, lo, hi)
    q = RU.linear_quantize(x, s, z, inplace=False)
    h = 2 ** (k - 1)
    return torch.clamp(q, -h, h - 1), s, z


def ref_codes_sym(x, k, lo, hi):
    s, z = RU.symmetric_linear_quantization_params_DSG(k, lo, hi)
    q = RU.linear_quantize_DSG(x, s, z, inplace=False)
    h = 2 ** (k - 1)
    return torch.clamp(q, -h, h - 1), s, z


# ------------------------------------------------------------------ frozen activations
def gen_act_frozen():
    g = torch.Generator().manual_seed(1234)
    out = {}
    shapes = {"hw49": (4, 8, 7, 7), "hw16": (3, 5, 4, 4), "flat": (2, 3, 16, 16)}
    for tag, shp in shapes.items():
        for k in (2, 3, 4, 8):
            x = torch.relu(torch.randn(shp, generator=g) * 1.7)
            lo, hi = x.min().reshape(1), (x.max() * 0.83).reshape(1)   # clip some of the tail
            q, s, z = ref_codes(x, k, lo, hi)
            y = RU.AsymmetricQuantFunction.apply(x, k, lo, hi)
            p = f"{tag}_k{k}_"
            out.update({p + "x": npy(x), p + "lo": npy(lo), p + "hi": npy(hi), p + "codes": npy(q),
                        p + "y": npy(y), p + "scale": npy(s), p + "zp": npy(z)})
    # signed input with a negative lower bound, and rounding ties
    for k in (2, 4, 8):
        x = torch.randn((2, 6, 5, 5), generator=g) * 2.0
        lo, hi = torch.tensor([-1.5]), torch.tensor([2.25])
        q, s, z = ref_codes(x, k, lo, hi)
        y = RU.AsymmetricQuantFunction.apply(x, k, lo, hi)
        p = f"signed_k{k}_"
        out.update({p + "x": npy(x), p + "lo": npy(lo), p + "hi": npy(hi), p + "codes": npy(q),
                    p + "y": npy(y), p + "scale": npy(s), p + "zp": npy(z)})
    # exact .5 ties: scale = 1 (range == n) so every x + 0.5 is a tie
    k = 4
    x = (torch.arange(-12, 12, dtype=torch.float32) + 0.5).reshape(1, 2, 3, 4)
    lo, hi = torch.tensor([-8.0]), torch.tensor([7.0])
    q, s, z = ref_codes(x, k, lo, hi)
    y = RU.AsymmetricQuantFunction.apply(x, k, lo, hi)
    out.update({"ties_k4_x": npy(x), "ties_k4_lo": npy(lo), "ties_k4_hi": npy(hi), "ties_k4_codes": npy(q),
                "ties_k4_y": npy(y), "ties_k4_scale": npy(s), "ties_k4_zp": npy(z)})
    # degenerate range (never-calibrated QuantAct: x_min == x_max == 0)
    x = torch.relu(torch.randn((2, 3, 4, 4), generator=g))
    lo, hi = torch.zeros(1), torch.zeros(1)
    q, s, z = ref_codes(x, k, lo, hi)
    y = RU.AsymmetricQuantFunction.apply(x, k, lo, hi)
    out.update({"degen_k4_x": npy(x), "degen_k4_lo": npy(lo), "degen_k4_hi": npy(hi), "degen_k4_codes": npy(q),
                "degen_k4_y": npy(y), "degen_k4_scale": npy(s), "degen_k4_zp": npy(z)})
    save("act_frozen", **out)


# ------------------------------------------------------------------ calibrating sequences
def gen_act_calib():
    out = {}
    for cls_name, tag in (("QuantAct", "asym"), ("QuantAct_DSG", "sym")):
        for k in (2, 4, 8):
            g = torch.Generator().manual_seed(77 + k)
            m = getattr(RM, cls_name)(activation_bit=k)
            p = f"{tag}_k{k}_"
            for step in range(6):
                x = torch.randn((3, 5, 7, 7), generator=g) * (1.0 + 0.3 * step)
                if tag == "asym":
                    x = torch.relu(x)
                if step == 4:
                    m.fix()          # frozen forward in the middle of the sequence
                if step == 5:
                    m.unfix()
                y = m(x)
                out[p + f"x{step}"] = npy(x)
                out[p + f"y{step}"] = npy(y)
                out[p + f"state{step}"] = np.array(
                    [m.x_min.item(), m.x_max.item(), m.beta_t.item()], dtype=np.float32)
                out[p + f"state_bits{step}"] = np.concatenate(
                    [npy(m.x_min), npy(m.x_max), npy(m.beta_t)]).view(np.int32)
            out[p + "beta"] = npy(m.beta)
    # full_precision_flag: returns the input, still tracks the range
    m = RM.QuantAct(activation_bit=4, full_precision_flag=True)
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn((2, 4, 4, 4), generator=g))
    y = m(x)
    assert y is x
    out["fp_x"] = npy(x)
    out["fp_state"] = np.array([m.x_min.item(), m.x_max.item(), m.beta_t.item()], dtype=np.float32)
    save("act_calib", **out)


# ------------------------------------------------------------------ weights
def gen_weights():
    out = {}
    g = torch.Generator().manual_seed(4321)
    convs = {"c3x3": (16, 3, 3, 3), "c1x1": (8, 16, 1, 1), "c7x7": (4, 3, 7, 7), "wide": (5, 32, 3, 3)}
    for tag, shp in convs.items():
        for k in (2, 4, 8):
            for sym in (False, True):
                conv = nn.Conv2d(shp[1], shp[0], shp[2], padding=shp[2] // 2, bias=(tag == "c1x1"))
                with torch.no_grad():
                    conv.weight.copy_(torch.randn(shp, generator=g) * 0.05)
                cls = RM.QuantConv2d_DSG if sym else RM.Quant_Conv2d
                qm = cls(weight_bit=k)
                qm.set_param(conv)
                x = torch.randn((2, shp[1], 6, 6), generator=g)
                yy = qm(x)
                yy.square().sum().backward()
                # the quantised weight itself, through the same calls forward() makes
                rows = qm.weight.data.contiguous().view(qm.out_channels, -1)
                if sym:
                    lo, hi = -rows.abs().max(dim=1).values, rows.abs().max(dim=1).values
                    wq = RU.SymmetricQuantFunction_DSG.apply(qm.weight.data, k, lo, hi)
                    q, s, z = ref_codes_sym(qm.weight.data, k, lo, hi)
                else:
                    lo, hi = rows.min(dim=1).values, rows.max(dim=1).values
                    wq = RU.AsymmetricQuantFunction.apply(qm.weight.data, k, lo, hi)
                    q, s, z = ref_codes(qm.weight.data, k, lo, hi)
                p = f"{tag}_k{k}_{'sym' if sym else 'asym'}_"
                out.update({p + "w": npy(qm.weight), p + "wq": npy(wq), p + "codes": npy(q), p + "lo": npy(lo),
                            p + "hi": npy(hi), p + "scale": npy(s), p + "zp": npy(z), p + "x": npy(x),
                            p + "out": npy(yy), p + "wgrad": npy(qm.weight.grad)})
                if qm.bias is not None:
                    out[p + "bias"] = npy(qm.bias)
    for k in (2, 4, 8):
        for sym in (False, True):
            lin = nn.Linear(64, 10)
            with torch.no_grad():
                lin.weight.copy_(torch.randn((10, 64), generator=g) * 0.1)
                lin.bias.copy_(torch.randn((10,), generator=g) * 0.1)
            cls = RM.QuantLinear_DSG if sym else RM.Quant_Linear
            qm = cls(weight_bit=k)
            qm.set_param(lin)
            x = torch.randn((7, 64), generator=g)
            yy = qm(x)
            yy.square().sum().backward()
            p = f"lin_k{k}_{'sym' if sym else 'asym'}_"
            out.update({p + "w": npy(qm.weight), p + "bias": npy(qm.bias), p + "x": npy(x), p + "out": npy(yy),
                        p + "wgrad": npy(qm.weight.grad)})
    # a constant row (min == max) and a row containing exact zeros only
    w = torch.randn((4, 2, 3, 3), generator=g) * 0.1
    w[1] = 0.25
    w[2] = 0.0
    rows = w.view(4, -1)
    lo, hi = rows.min(dim=1).values, rows.max(dim=1).values
    wq = RU.AsymmetricQuantFunction.apply(w, 4, lo, hi)
    out.update({"constrow_w": npy(w), "constrow_wq": npy(wq)})
    save("weights", **out)


# ------------------------------------------------------------------ MSE-searched activation range
def gen_act_mse():
    out = {}
    g = torch.Generator().manual_seed(99)
    m = RM.QuantAct_MSE(activation_bit=4)
    for step in range(2):
        x = torch.relu(torch.randn((2, 4, 6, 6), generator=g) * 1.5)
        y = m(x)
        out[f"x{step}"] = npy(x)
        out[f"y{step}"] = npy(y)
        out[f"state{step}"] = np.array([m.x_min.item(), m.x_max.item(), m.beta_t.item()], dtype=np.float32)
    save("act_mse", **out)


# ------------------------------------------------------------------ BN-statistics loss
class TinyNet(nn.Module):
    """conv-bn-relu x3 with planes of 49 / 16 elements; only a carrier for BN hooks."""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 6, 3, padding=1, bias=False)
        self.b1 = nn.BatchNorm2d(6)
        self.c2 = nn.Conv2d(6, 10, 3, stride=2, padding=1, bias=False)
        self.b2 = nn.BatchNorm2d(10)
        self.c3 = nn.Conv2d(10, 4, 1, bias=False)
        self.b3 = nn.BatchNorm2d(4)

    def forward(self, x):
        x = torch.relu(self.b1(self.c1(x)))
        x = torch.relu(self.b2(self.c2(x)))
        return self.b3(self.c3(x))


def gen_bns():
    torch.manual_seed(11)
    net = TinyNet().eval()
    g = torch.Generator().manual_seed(12)
    for bn in (net.b1, net.b2, net.b3):
        bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
    dd = DistillData()                                   # the reference hook, verbatim
    for bn in (net.b1, net.b2, net.b3):
        bn.register_forward_hook(dd.hook_fn_forward)
    out = {f"param_{n}": npy(p) for n, p in net.state_dict().items()}
    mse = nn.MSELoss()
    for flavour in ("trainer", "distill"):
        x = (torch.randn((5, 3, 14, 14), generator=torch.Generator().manual_seed(13)) * 0.7 + 0.2).requires_grad_(True)
        for lst in (dd.mean_list, dd.var_list, dd.teacher_running_mean, dd.teacher_running_var):
            lst.clear()
        net(x)
        L = len(dd.mean_list)
        if flavour == "trainer":      # same calls as trainer_direct.py:474-484 (file itself cannot be imported)
            loss = torch.zeros(1)
            for i in range(L):
                loss += mse(dd.mean_list[i], dd.teacher_running_mean[i]) + mse(dd.var_list[i], dd.teacher_running_var[i])
            loss = loss / L
        else:                         # same calls as distill_data.py:252-265
            ml, vl = torch.zeros(1), torch.zeros(1)
            for i in range(L):
                ml += mse(dd.mean_list[i], dd.teacher_running_mean[i].detach())
                vl += mse(dd.var_list[i], dd.teacher_running_var[i].detach())
            loss = ml / L + vl / L
        loss.backward()
        out[f"{flavour}_loss"] = npy(loss)
        out[f"{flavour}_xgrad"] = npy(x.grad)
        out["x"] = npy(x)
        for i in range(L):
            out[f"mean{i}"] = npy(dd.mean_list[i])
            out[f"var{i}"] = npy(dd.var_list[i])
    # raw statistics of tensors with a large mean/sigma ratio (cancellation stress)
    g = torch.Generator().manual_seed(14)
    for tag, off in (("off0", 0.0), ("off10", 10.0), ("off100", 100.0)):
        t = torch.randn((6, 5, 7, 7), generator=g) + off
        out[f"stat_{tag}_x"] = npy(t)
        out[f"stat_{tag}_mean"] = npy(t.mean([0, 2, 3]))
        out[f"stat_{tag}_var"] = npy(t.var([0, 2, 3], unbiased=False))
        out[f"stat_{tag}_var64"] = npy(t.double().var([0, 2, 3], unbiased=False))
    save("bns", **out)


# ------------------------------------------------------------------ per-sample augmentation (torchvision)
def gen_augment():
    """The transform objects of direct_dataset (main_direct.py:158-169), built exactly as there, run per sample at
    fixed seeds.  The pipeline hides its draws, so each sample is run twice from the same generator state: once
    through the Compose itself (the expected output) and once through torchvision's own get_params / the flip's
    ``torch.rand(1) < p`` (the recorded box and flip); the decomposition is asserted to reproduce the Compose
    bit for bit before anything is saved.  ``*_out_exact`` comes from the same torchvision calls on the image in
    double precision (coordinates and weights then carry no fp32 rounding), rounded to fp32 once at the end."""
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    out = {"torchvision_version": np.array(__import__("torchvision").__version__)}
    sets = {"rgb32": (4, 3, 32, 32), "grey28": (5, 1, 28, 28), "tall160": (2, 3, 160, 16), "rect": (3, 3, 20, 36),
            "big200": (1, 1, 200, 200)}   # small files
    for tag, shape in sets.items():
        g = torch.Generator().manual_seed(sum(shape))
        images = torch.randn(shape, generator=g)
        size = shape[2] if shape[2] == shape[3] else (shape[2], shape[3])
        pipeline = T.Compose([
            T.RandomResizedCrop(size=size, scale=(0.5, 1.0)),
            T.Lambda(lambda x: x.repeat(3, 1, 1) if x.size(0) == 1 else x),
            T.RandomHorizontalFlip(),
        ])
        crop = pipeline.transforms[0]
        samples = 2 * shape[0] + 1 if tag != "big200" else 2
        index = torch.randint(0, shape[0], (samples,), generator=g)
        boxes, flips, ys, ys64 = [], [], [], []
        torch.manual_seed(77 + shape[2])
        for m in index.tolist():
            state = torch.get_rng_state()
            y = pipeline(images[m])
            after = torch.get_rng_state()
            torch.set_rng_state(state)
            i, j, h, w = T.RandomResizedCrop.get_params(images[m], crop.scale, crop.ratio)
            flip = bool(torch.rand(1) < 0.5)
            assert torch.equal(torch.get_rng_state(), after)
            z = TF.resized_crop(images[m], i, j, h, w, crop.size, crop.interpolation, antialias=crop.antialias)
            z = z.repeat(3, 1, 1) if z.size(0) == 1 else z
            z = TF.hflip(z) if flip else z
            assert torch.equal(z, y), "decomposition differs from the Compose pipeline"
            z64 = TF.resized_crop(images[m].double(), i, j, h, w, crop.size, crop.interpolation, antialias=crop.antialias)
            z64 = z64.repeat(3, 1, 1) if z64.size(0) == 1 else z64
            z64 = TF.hflip(z64) if flip else z64
            # for an up-scaling the antialiased filter is the plain bilinear one
            plain = TF.resized_crop(images[m].double(), i, j, h, w, crop.size, crop.interpolation, antialias=False)
            plain = plain.repeat(3, 1, 1) if plain.size(0) == 1 else plain
            assert (z64 - (TF.hflip(plain) if flip else plain)).abs().max() < 1e-12
            boxes.append([i, j, h, w]); flips.append(flip); ys.append(y); ys64.append(z64)
        # gradient reaching the image set through torchvision's own (differentiable) pipeline, as autograd computes it
        # for the distillation loop's RHF(RRC(gaussian_data[j])) (data_generate/distill_data.py:197-227): a fixed
        # cotangent per sample, float64 so that the result carries no accumulation-order noise
        leaf = images.double().requires_grad_(True)
        cot = torch.randn((samples,) + tuple(ys[0].shape), generator=g)
        if tag == "big200":                                        # small file: one cotangent plane for all three channels
            cot = cot[:, :1].repeat(1, 3, 1, 1)
        total = 0.0
        for s_i, m in enumerate(index.tolist()):
            i, j, h, w = boxes[s_i]
            z = TF.resized_crop(leaf[m], i, j, h, w, crop.size, crop.interpolation, antialias=crop.antialias)
            z = z.repeat(3, 1, 1) if z.size(0) == 1 else z
            z = TF.hflip(z) if flips[s_i] else z
            total = total + (z * cot[s_i].double()).sum()
        total.backward()
        out[f"{tag}_cotangent"] = npy(cot if tag != "big200" else cot[:, :1])
        out[f"{tag}_grad_exact"] = leaf.grad.float().numpy().copy()
        out[f"{tag}_images"] = npy(images)
        out[f"{tag}_index"] = index.numpy().astype(np.int64)
        out[f"{tag}_boxes"] = np.array(boxes, dtype=np.int32)
        out[f"{tag}_flips"] = np.array(flips, dtype=np.uint8)
        out[f"{tag}_seed"] = np.array(77 + shape[2])
        if tag == "big200":          # coordinates up to 200; the three repeated channels are identical: keep one
            assert all(torch.equal(y[0], y[1]) and torch.equal(y[0], y[2]) for y in ys)
            ys, ys64 = [y[:1] for y in ys], [y[:1] for y in ys64]
        out[f"{tag}_out"] = npy(torch.stack(ys))
        out[f"{tag}_out_exact"] = torch.stack(ys64).float().numpy().copy()     # double result, rounded once
    save("augment", **out)


if __name__ == "__main__":
    gen_act_frozen()
    gen_act_calib()
    gen_weights()
    gen_act_mse()
    gen_bns()
    gen_augment()
