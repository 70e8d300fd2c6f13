"""The three arithmetic identities the table kernels rely on (csrc/common.cuh ``fake_quant_lut`` /
``relu_fake_quant_lut``, csrc/bn_pool_ring.cuh packed candidates), restated in numpy float32 and checked against the
reference's order of operations (quant_utils.py:150-152: ``round(scale*x - zp)`` then ``clamp``) on random and
adversarial inputs.  CPU only; the kernels themselves are compared bit for bit with the reference chain by the -m gpu
suite."""
import numpy as np
import pytest

F = np.float32
MAGIC = F(12582912.0)          # 1.5 * 2^23


def params(k, lo, hi):
    """make_qparams (quant_utils.py:117-128): reciprocal * n, rint, + 2^(k-1)."""
    r = np.maximum(F(hi) - F(lo), F(1e-8))
    scale = (F(1.0) / r) * F(2 ** k - 1)
    zp = np.rint(scale * F(lo)).astype(F) + F(2 ** (k - 1))
    return scale.astype(F), zp.astype(F), F(-2 ** (k - 1)), F(2 ** (k - 1) - 1)


def reference_code(x, scale, zp, qlo, qhi):
    u = (scale * x).astype(F) - zp                     # two fp32 roundings
    return np.clip(np.rint(u.astype(F)), qlo, qhi).astype(F)


def samples(rng, scale, zp, n=200000):
    x = (rng.standard_normal(n) * 2.0).astype(F)
    # every half-way point of the quantiser, the points around them, zeros, infinities and huge values
    codes = np.arange(-300, 300, dtype=np.float64)
    for d in (-0.5, 0.5, 0.0):
        pts = ((codes + d + float(zp)) / float(scale)).astype(F)
        x = np.concatenate([x, pts, np.nextafter(pts, F(np.inf)), np.nextafter(pts, F(-np.inf))])
    return np.concatenate([x, np.array([0.0, -0.0, np.inf, -np.inf, 3e38, -3e38, 1e-45, -1e-45], dtype=F)])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("lo,hi", [(0.0, 1.9), (-1.3, 2.6), (0.0, 1e-9), (-5.0, -1.0), (8388608.0, 8388612.0)])
def test_clamp_first_then_magic_add_is_round_then_clamp(k, lo, hi):
    """code = clamp(rint(u)) = low mantissa bits of (clamp(u) + 1.5*2^23)."""
    rng = np.random.default_rng(k)
    scale, zp, qlo, qhi = params(k, lo, hi)
    x = samples(rng, scale, zp)
    with np.errstate(over="ignore", invalid="ignore"):
        ref = reference_code(x, scale, zp, qlo, qhi)
        u = (scale * x).astype(F) - zp
        key = (np.clip(u.astype(F), qlo, qhi) + MAGIC).astype(F)
    code = (key.view(np.int32) - MAGIC.view(np.int32)).astype(F)
    ok = ~np.isnan(ref)
    assert np.array_equal(code[ok], ref[ok])
    h, mask = 2 ** (k - 1), 2 ** k - 1
    assert np.array_equal((key.view(np.int32)[ok] + h) & mask, (ref[ok].astype(np.int64) + h) & mask)     # the table index


@pytest.mark.parametrize("k", [1, 2, 4, 8])
@pytest.mark.parametrize("lo,hi", [(0.0, 1.9), (-1.3, 2.6), (0.3, 2.0), (8388608.0, 8388612.0)])
def test_relu_is_the_lower_clamp(k, lo, hi):
    """code(relu(z)) = min(max(rint(u(z)), max(-zp, qlo)), qhi): u is monotone, u(+-0) = -zp, zp is an integer."""
    rng = np.random.default_rng(10 + k)
    scale, zp, qlo, qhi = params(k, lo, hi)
    assert float(zp) == np.rint(float(zp))
    z = samples(rng, scale, zp)
    with np.errstate(over="ignore", invalid="ignore"):
        ref = reference_code(np.maximum(z, F(0.0)), scale, zp, qlo, qhi)
        u = (scale * z).astype(F) - zp
        lowc = np.maximum(-zp, qlo)
        key = (np.minimum(np.maximum(u.astype(F), lowc), qhi) + MAGIC).astype(F)
    code = (key.view(np.int32) - MAGIC.view(np.int32)).astype(F)
    assert np.array_equal(code, ref)


def test_packed_candidates_pick_atens_first_maximum():
    """(bits(code + 1.5*2^23) << 4) + tie-break, unsigned max over the window = max_pool2d's first maximum in scan order;
    the tie-break 4*(2-r) + (2-j) decodes to the window-local index 3r + j."""
    rng = np.random.default_rng(3)
    k = 4
    scale, zp, qlo, qhi = params(k, 0.0, 1.9)
    wins = (rng.standard_normal((20000, 3, 3)) * 1.2).astype(F)
    wins[:2000] = np.round(wins[:2000])                 # many ties
    wins[2000:2100] = F(0.25)                           # nothing but ties
    valid = rng.random((20000, 3, 3)) < 0.9             # border windows: some positions are outside the image
    valid[:, 1, 1] = True
    code = reference_code(np.maximum(wins, F(0.0)), scale, zp, qlo, qhi)
    key = (code + MAGIC).astype(F)
    tb = np.array([[4 * (2 - r) + (2 - j) for j in range(3)] for r in range(3)], dtype=np.uint32)
    packed = (key.view(np.uint32) << np.uint32(4)) + tb[None]
    packed = np.where(valid, packed, np.uint32(0))
    best = packed.reshape(-1, 9).max(axis=1)
    t = best & np.uint32(15)
    pos = 3 * (2 - (t >> np.uint32(2))) + (2 - (t & np.uint32(3)))
    got_code = ((best >> np.uint32(4)).astype(np.int64) - (int(MAGIC.view(np.uint32)) & 0x0FFFFFFF)).astype(F)
    # ATen: scan the window row-major, a later element replaces the running maximum only when it is greater
    flat_c = np.where(valid, code, F(-np.inf)).reshape(-1, 9)
    ref_pos = flat_c.argmax(axis=1)                     # numpy's argmax returns the FIRST maximum
    assert np.array_equal(pos.astype(np.int64), ref_pos)
    assert np.array_equal(got_code, flat_c.max(axis=1))
    # a NaN key shifted left by four exceeds every regular packed value (the kernel's slow-path trigger)
    nan_key = (np.array([np.nan], dtype=F) + MAGIC).astype(F).view(np.uint32) << np.uint32(4)
    assert nan_key[0] >= np.uint32(0xC0000000) > packed.max()
