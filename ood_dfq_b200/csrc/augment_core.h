// Per-pixel body of the crop -> bilinear resize -> horizontal flip kernel (augment.cu).
//
// Kept in a header of plain functions so that the very same index arithmetic and blend that the sm_100a kernel
// runs can also be compiled by g++ inside tests/test_augment_cpu.py and checked against the oracle without a
// GPU.  That host build is test scaffolding only: nothing in the package can reach it, and ops.crop_resize_flip
// raises for CPU tensors like every other entry point.
#pragma once

#if defined(__CUDACC__)
#define OODFQ_HD __host__ __device__ __forceinline__
#else
#define OODFQ_HD static inline
#endif

namespace oodfq {

struct AugGeom {
    long long M;             // stored images [M, C_in, H, W], NCHW as the shards hold them
    int C_in, H, W;
    int N, OH, OW;           // batch of N outputs [N, C_out, OH, OW]
    int nhwc;                // output is channels_last ([N, OH, OW, C_out])
    int src_nhwc;            // stored images are channels_last ([M, H, W, C_in]) instead of NCHW
};

struct AugArgs {
    const float* images;           // device-resident image set (backward: unused)
    const long long* index;        // [N] sample -> image
    const int* boxes;              // [N][4] (top, left, height, width) of the crop, inside the image
    const unsigned char* flips;    // [N] non-zero: mirror the resized crop left-right
    float* out;                    // forward: the batch; backward: the gradient w.r.t. the image set (accumulated)
    const float* grad_out;         // backward only: gradient w.r.t. the batch, in the batch's layout
    AugGeom g;
};

struct AugTap {
    int i0, i1;
    float l0, l1;
};

// Source taps of output coordinate `dst` when `in` source pixels are resized to `out` (align_corners=False):
// src = (in/out)(dst + 1/2) - 1/2 clamped at 0, i0 = floor(src), i1 = min(i0 + 1, in - 1), l1 = src - i0 --
// ATen's area_pixel_compute_source_index / upsample_bilinear2d.  The reference evaluates src in fp32, whose
// spacing at coordinate 200 is 1.5e-5; here src is the exact rational (in(2 dst + 1) - out) / (2 out) and only
// the final weight is rounded, so the result is the same on every compiler (no FMA ambiguity) and at least as
// close to the real-valued bilinear filter as the reference's own.
// Sizes are at most kAugMaxSide, so every product below fits 32 bits (64-bit division is a subroutine on the GPU).
constexpr int kAugMaxSide = 16384;

OODFQ_HD AugTap aug_tap(int dst, int in, int out) {
    AugTap t;
    const int num = in * (2 * dst + 1) - out;
    const int den = 2 * out;
    if (num <= 0) {
        t.i0 = 0;
        t.l1 = 0.0f;
    } else {
        int q = num / den;
        const int rem = num - q * den;
        t.l1 = (float)rem / (float)den;
        if (q > in - 1) {            // cannot happen for dst < out; keeps every read inside the box regardless
            q = in - 1;
            t.l1 = 0.0f;
        }
        t.i0 = q;
    }
    if (t.i0 < in - 1) {
        t.i1 = t.i0 + 1;
    } else {                         // last source pixel: both taps are the same pixel, weight it once (ATen blends
        t.i1 = t.i0;                 // p*l0 + p*l1, which can be an ulp off p)
        t.l1 = 0.0f;
    }
    t.l0 = 1.0f - t.l1;
    return t;
}

#if defined(__CUDA_ARCH__)
#define OODFQ_AUG_LOAD(p) __ldg(p)
#else
#define OODFQ_AUG_LOAD(p) (*(p))
#endif

// Where one output pixel reads from: image, the two source rows / columns (offsets inside a plane) and weights.
struct AugSite {
    int n, oy, ox;
    long long im;                  // image index inside the set
    int y0, y1, x0, x1;            // absolute source rows / columns
    float wy0, wy1, wx0, wx1;
};

// `pix` counts (n, oy, ox) row-major (< 2^31 per launch).
OODFQ_HD AugSite aug_site(const AugArgs& a, int pix) {
    const AugGeom& g = a.g;
    AugSite s;
    const int q = pix / g.OW;
    s.ox = pix - q * g.OW;
    s.n = q / g.OH;
    s.oy = q - s.n * g.OH;
    // boxes and indices come from device memory the host side cannot re-check at launch time: fold them into the
    // image set so that a corrupt entry can never turn into an out-of-bounds access (valid entries are unchanged)
    int bh = a.boxes[4 * s.n + 2], bw = a.boxes[4 * s.n + 3];
    bh = bh < 1 ? 1 : (bh > g.H ? g.H : bh);
    bw = bw < 1 ? 1 : (bw > g.W ? g.W : bw);
    int top = a.boxes[4 * s.n], left = a.boxes[4 * s.n + 1];
    top = top < 0 ? 0 : (top > g.H - bh ? g.H - bh : top);
    left = left < 0 ? 0 : (left > g.W - bw ? g.W - bw : left);
    long long im = a.index[s.n];
    s.im = im < 0 ? 0 : (im > g.M - 1 ? g.M - 1 : im);
    // RandomHorizontalFlip runs AFTER the resize (main_direct.py:160-162): output column ox shows resized column
    // OW-1-ox
    const int rx = a.flips[s.n] ? g.OW - 1 - s.ox : s.ox;
    const AugTap ty = aug_tap(s.oy, bh, g.OH), tx = aug_tap(rx, bw, g.OW);
    s.y0 = top + ty.i0; s.y1 = top + ty.i1; s.x0 = left + tx.i0; s.x1 = left + tx.i1;
    s.wy0 = ty.l0; s.wy1 = ty.l1; s.wx0 = tx.l0; s.wx1 = tx.l1;
    return s;
}

// element offset of (image, channel, row, column) in the stored set, either layout
template <int C_IN>
OODFQ_HD long long aug_src_offset(const AugGeom& g, long long im, int c, int y, int x) {
    return g.src_nhwc ? ((im * g.H + y) * g.W + x) * C_IN + c : ((im * C_IN + c) * g.H + y) * g.W + x;
}

// One output pixel, every stored channel: v[c] for c < C_IN.
template <int C_IN>
OODFQ_HD void aug_pixel(const AugArgs& a, int pix, float* v, int& n, int& oy, int& ox) {
    const AugSite s = aug_site(a, pix);
    n = s.n; oy = s.oy; ox = s.ox;
#pragma unroll
    for (int c = 0; c < C_IN; ++c) {
        const float p00 = OODFQ_AUG_LOAD(a.images + aug_src_offset<C_IN>(a.g, s.im, c, s.y0, s.x0));
        const float p01 = OODFQ_AUG_LOAD(a.images + aug_src_offset<C_IN>(a.g, s.im, c, s.y0, s.x1));
        const float p10 = OODFQ_AUG_LOAD(a.images + aug_src_offset<C_IN>(a.g, s.im, c, s.y1, s.x0));
        const float p11 = OODFQ_AUG_LOAD(a.images + aug_src_offset<C_IN>(a.g, s.im, c, s.y1, s.x1));
        v[c] = s.wy0 * (s.wx0 * p00 + s.wx1 * p01) + s.wy1 * (s.wx0 * p10 + s.wx1 * p11);
    }
}

#if defined(__CUDA_ARCH__)
#define OODFQ_AUG_ADD(p, v) atomicAdd((p), (v))
#else
#define OODFQ_AUG_ADD(p, v) (*(p) += (v))
#endif

// Backward of one output pixel: its gradient (summed over the repeated channels of a grey image) is scattered onto
// the four taps with the forward's weights.  Taps are shared between neighbouring outputs, hence atomic adds on the
// device (ATen's upsample backward does the same); zero-weight taps are skipped.
template <int C_IN, int C_OUT>
OODFQ_HD void aug_pixel_backward(const AugArgs& a, int pix) {
    const AugGeom& g = a.g;
    const AugSite s = aug_site(a, pix);
    const long long oplane = (long long)g.OH * g.OW;
#pragma unroll
    for (int c = 0; c < C_IN; ++c) {
        float go = 0.0f;
#pragma unroll
        for (int r = (C_IN == C_OUT ? c : 0); r < (C_IN == C_OUT ? c + 1 : C_OUT); ++r)
            go += OODFQ_AUG_LOAD(a.grad_out + (g.nhwc ? (long long)pix * C_OUT + r
                                                      : ((long long)s.n * C_OUT + r) * oplane + (long long)s.oy * g.OW + s.ox));
        const float w00 = s.wy0 * s.wx0, w01 = s.wy0 * s.wx1, w10 = s.wy1 * s.wx0, w11 = s.wy1 * s.wx1;
        if (w00 != 0.0f) OODFQ_AUG_ADD(a.out + aug_src_offset<C_IN>(g, s.im, c, s.y0, s.x0), w00 * go);
        if (w01 != 0.0f) OODFQ_AUG_ADD(a.out + aug_src_offset<C_IN>(g, s.im, c, s.y0, s.x1), w01 * go);
        if (w10 != 0.0f) OODFQ_AUG_ADD(a.out + aug_src_offset<C_IN>(g, s.im, c, s.y1, s.x0), w10 * go);
        if (w11 != 0.0f) OODFQ_AUG_ADD(a.out + aug_src_offset<C_IN>(g, s.im, c, s.y1, s.x1), w11 * go);
    }
}

// PX consecutive output pixels starting at pixel `first`, stored in the layout the geometry names.  A
// one-channel image feeding a three-channel output is repeated (the Lambda of main_direct.py:161).
// channels_last with PX = 4: the PX * C_OUT floats of the group are contiguous and 16-byte aligned -> float4 stores.
template <int C_IN, int C_OUT, int PX>
OODFQ_HD void aug_group(const AugArgs& a, int first, int total) {
    const AugGeom& g = a.g;
    float buf[PX * C_OUT];
    int n[PX], oy[PX], ox[PX];
    const int cnt = (total - first) < PX ? (total - first) : PX;
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        if (p < cnt) {
            float v[C_IN];
            aug_pixel<C_IN>(a, first + p, v, n[p], oy[p], ox[p]);
#pragma unroll
            for (int c = 0; c < C_OUT; ++c) buf[p * C_OUT + c] = v[C_IN == C_OUT ? c : 0];
        }
    }
    if (g.nhwc) {
        float* o = a.out + (long long)first * C_OUT;
#if defined(__CUDA_ARCH__)
        if (PX == 4 && cnt == PX) {
#pragma unroll
            for (int j = 0; j < (PX * C_OUT) / 4; ++j)
                reinterpret_cast<float4*>(o)[j] = make_float4(buf[4 * j], buf[4 * j + 1], buf[4 * j + 2], buf[4 * j + 3]);
            return;
        }
#endif
#pragma unroll
        for (int p = 0; p < PX; ++p)
            if (p < cnt) {
#pragma unroll
                for (int c = 0; c < C_OUT; ++c) o[p * C_OUT + c] = buf[p * C_OUT + c];
            }
    } else {
        const long long oplane = (long long)g.OH * g.OW;
#pragma unroll
        for (int p = 0; p < PX; ++p)
            if (p < cnt) {
#pragma unroll
                for (int c = 0; c < C_OUT; ++c)
                    a.out[((long long)n[p] * C_OUT + c) * oplane + (long long)oy[p] * g.OW + ox[p]] = buf[p * C_OUT + c];
            }
    }
}

}  // namespace oodfq
