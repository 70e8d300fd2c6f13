// Stem fusion: eval-mode BatchNorm -> ReLU -> [QuantAct] -> MaxPool2d(3, stride 2, padding 1) as one kernel
// forward and one kernel backward, channels_last.
//
// In the ImageNet ResNets the first QuantAct site is followed directly by the stem max-pool
// (pytorchcv ResInitBlock: conv7x7 block -> MaxPool2d(3,2,1); reference models via ptcv_get_model,
// main_direct.py:380-397).  After the BatchNorm fusion, ATen's max_pool forward+backward were 16 % of the
// step (profiles/r1_step_share_final.txt): they move int64 argmax indices (1.6 GB per launch) and the
// full-resolution 822 MB tensor once more.  Here the full-resolution quantised tensor never exists:
//
//   forward   out[n,ho,wo,c] = max over the 3x3 window of  y = fakequant(relu(a_c*x + b_c))
//             (first maximum in window scan order, exactly what max_pool2d of the unfused chain picks);
//             reads x once (re-reads of neighbouring windows hit L1/L2), writes out (1/4 size), a one-byte
//             argmax code per output (window-local index + ReLU-active bit) and, when BatchNorm parameter
//             gradients are wanted, the normalised input at the argmax.
//   backward  grad_x[n,h,w,c] = a_c * sum over the <= 4 windows covering (h,w) whose argmax is (h,w) and whose
//             ReLU was active of grad_out (gather, fixed order: deterministic, no atomics); x is not read at
//             all.  dB_c, dW_c are accumulated by the thread sitting on each window's centre.
//
// Traffic for [256,64,112,112]: forward 822 MB read + 0.46 GB written, backward 0.46 GB read + 822 MB written,
// against 1.6 + 2.7 GB and 2.5 + 2.7 GB for fused-BN followed by ATen max-pool.  Roofline: HBM.
#include "bn_geom.cuh"

namespace oodfq {

struct BnParams2 {
    const float* w;
    const float* b;
    const float* rm;
    const float* rv;
    float eps;
};

__device__ __forceinline__ void affine2(const BnParams2& P, int c, float& a, float& b, float& invstd) {
    invstd = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(P.rv + c), P.eps)));
    a = __fmul_rn(P.w ? __ldg(P.w + c) : 1.0f, invstd);
    b = __fsub_rn(P.b ? __ldg(P.b + c) : 0.0f, __fmul_rn(__ldg(P.rm + c), a));
}

struct PoolGeom {
    int N, C, H, W, Ho, Wo, cols, lanes_r;
    int seg, nseg;          // forward: output rows per work item, work items per column of windows
};

// One candidate of the running "first maximum in window scan order" search, per channel.
//   key   what max_pool2d compares: the fake-quantised value's integer code when the dequantisation table is
//         strictly increasing (the code orders exactly like the value, and skips the table), else the value
//   meta  column of the candidate inside its window row (0..2) | 128 if the ReLU was active there
//   x     the raw input at the candidate (only kept when the normalised input is wanted)
struct Cand {
    float key[4];
    int meta[4];
    float x[4];
};

// ATen's max_pool2d rule: a later element replaces the running maximum when it is greater or NaN
__device__ __forceinline__ bool takes_over(float v, float best) { return v > best || v != v; }

template <bool QUANT, bool STRICT, bool XHAT>
__device__ __forceinline__ void scan_pixel(Cand& r, const float4 v, const int j, const float (&a)[4], const float (&b)[4],
                                           const QParams& qp, const float* lut, const int qh, const int qmask) {
    const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float zr = fmaf(xs[c], a[c], b[c]);                     // BN affine
        float key;
        if (QUANT && STRICT) {
            // The code of relu(z) without the ReLU, the clamp-after-round and the FRND of code_of():
            //   u = scale*z - zp  (the reference's two roundings); relu: s*z - zp is monotone in z and equals -zp at z = 0
            //   (also at -0), so max(u, -zp) IS u(relu(z)); clamp and rint commute for integer bounds; rint of a value in
            //   [-2^22, 2^22] is (u + 1.5*2^23) - 1.5*2^23 in round-to-nearest-even.  The biased sum orders like the code
            //   and its low bits are the table index, so it is kept as the key: 6 FP32-pipe instructions, bit-identical
            //   codes (tests/test_gpu_fused.py: stem vs the unfused chain, NaN and merged tables included).
            float u = __fsub_rn(__fmul_rn(qp.scale, zr), qp.zp);
            u = min_nan(max_nan(max_nan(u, -qp.zp), qp.qlo), qp.qhi);
            key = __fadd_rn(u, kRoundMagic);
        } else {
            const float z = max_nan(zr, 0.0f);                         // ReLU (NaN stays NaN)
            key = z;
            if (QUANT) {
                key = code_of<false>(z, qp);
                const float y = lut[lut_index(key, qh, qmask)];
                key = (key != key) ? key : y;
            }
        }
        const bool t = takes_over(key, r.key[c]);
        r.key[c] = t ? key : r.key[c];
        r.meta[c] = t ? (zr > 0.0f ? (j | 128) : j) : r.meta[c];
        if (XHAT) r.x[c] = t ? xs[c] : r.x[c];
    }
}

__device__ __forceinline__ void reset(Cand& r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { r.key[c] = -INFINITY; r.meta[c] = 0; r.x[c] = 0.0f; }
}

// the three pixels of one input row inside the window columns 2*wo-1 .. 2*wo+1; `row` points at column 2*wo-1
// (this thread's four channels), `ok` says which of the three columns exist
struct RowPixels { float4 v[3]; };

__device__ __forceinline__ void load_row(RowPixels& p, const float4* __restrict__ row, const int cols, const bool (&ok)[3],
                                         const bool row_ok) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (row_ok && ok[j]) p.v[j] = __ldg(row + j * cols);
}

// best candidate of a loaded row
template <bool QUANT, bool STRICT, bool XHAT>
__device__ __forceinline__ void scan_row(Cand& r, const RowPixels& p, const bool (&ok)[3], const bool row_ok,
                                         const float (&a)[4], const float (&b)[4], const QParams& qp, const float* lut,
                                         const int qh, const int qmask) {
    reset(r);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (row_ok && ok[j]) scan_pixel<QUANT, STRICT, XHAT>(r, p.v[j], j, a, b, qp, lut, qh, qmask);
}

// A lane owns one column of windows (n, wo) and walks down a segment of output rows: window ho covers input
// rows 2ho-1, 2ho, 2ho+1, and the candidate of row 2ho+1 is carried over as row 2(ho+1)-1 of the next window,
// so every step evaluates 6 new pixels instead of 9.  The six loads of step ho+1 are issued before step ho is
// evaluated (register double buffer): at 2-3 resident CTAs per SM the kernel is otherwise latency-bound
// (ncu: 24 % warps active, long-scoreboard stalls; profiles/r1_bn_pool_kernels.txt).
template <bool QUANT, bool STRICT, bool XHAT>
__device__ __forceinline__ void pool_fwd_body(const float* __restrict__ x, float* __restrict__ out,
                                              uint8_t* __restrict__ idx, float* __restrict__ xhat, const PoolGeom& G,
                                              const BnParams2& P, const QParams& qp, const float* lut, const int qh,
                                              const int qmask) {
    if ((int)threadIdx.x >= G.lanes_r * G.cols) return;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a[4], b[4], rm[4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { affine2(P, 4 * col + j, a[j], b[j], inv[j]); rm[j] = __ldg(P.rm + 4 * col + j); }
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long columns = (long long)G.N * G.Wo;
    const long long groups = (columns + G.lanes_r - 1) / G.lanes_r;
    const long long items = groups * G.nseg;
    const long long row_stride = (long long)G.W * G.cols;            // float4 units
    const int ostride = G.Wo * G.cols;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long u = (item / G.nseg) * G.lanes_r + rsub;
        if (u >= columns) continue;
        const int sidx = (int)(item % G.nseg);
        const int wo = (int)(u % G.Wo);
        const long long n = u / G.Wo;
        const int ho_begin = sidx * G.seg, ho_end = min(G.Ho, ho_begin + G.seg);
        const bool ok[3] = {wo > 0, true, 2 * wo + 1 < G.W};
        // input row 2*ho_begin - 1, column 2*wo - 1 (either may lie outside the image: never dereferenced then)
        const float4* row = x4 + (n * G.H + (2 * ho_begin - 1)) * row_stride + (long long)(2 * wo - 1) * G.cols + col;
        long long o = ((n * G.Ho + ho_begin) * G.Wo + wo) * G.cols + col;
        Cand carried, r1, r2;
        RowPixels p0, p1, p2, q1, q2;
        load_row(p0, row, G.cols, ok, ho_begin > 0);
        load_row(p1, row + row_stride, G.cols, ok, true);
        load_row(p2, row + 2 * row_stride, G.cols, ok, 2 * ho_begin + 1 < G.H);
        scan_row<QUANT, STRICT, XHAT>(carried, p0, ok, ho_begin > 0, a, b, qp, lut, qh, qmask);
        for (int ho = ho_begin; ho < ho_end; ++ho, o += ostride) {
            row += 2 * row_stride;                       // now at input row 2*ho + 1
            const bool more = ho + 1 < ho_end;           // rows 2ho+2 (always inside the image then) and 2ho+3
            load_row(q1, row + row_stride, G.cols, ok, more);
            load_row(q2, row + 2 * row_stride, G.cols, ok, more && 2 * ho + 3 < G.H);
            scan_row<QUANT, STRICT, XHAT>(r1, p1, ok, true, a, b, qp, lut, qh, qmask);
            scan_row<QUANT, STRICT, XHAT>(r2, p2, ok, 2 * ho + 1 < G.H, a, b, qp, lut, qh, qmask);
            float y[4], xh[4];
            unsigned char code[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // rows in scan order; the window-local index is 3*row + column
                float key = carried.key[c], bx = carried.x[c];
                int meta = carried.meta[c];
                const bool t1 = takes_over(r1.key[c], key);
                key = t1 ? r1.key[c] : key; meta = t1 ? r1.meta[c] + 3 : meta; bx = t1 ? r1.x[c] : bx;
                const bool t2 = takes_over(r2.key[c], key);
                key = t2 ? r2.key[c] : key; meta = t2 ? r2.meta[c] + 6 : meta; bx = t2 ? r2.x[c] : bx;
                if (QUANT && STRICT) {
                    // key = code + 1.5*2^23 (see scan_pixel): its low bits index the table directly
                    const float v = lut[(__float_as_int(key) + qh) & qmask];
                    y[c] = (key != key) ? key : v;
                } else {
                    y[c] = key;
                }
                code[c] = (unsigned char)meta;
                xh[c] = (bx - rm[c]) * inv[c];
            }
            st_out(reinterpret_cast<float4*>(out) + o, make_float4(y[0], y[1], y[2], y[3]));
            reinterpret_cast<uchar4*>(idx)[o] = make_uchar4(code[0], code[1], code[2], code[3]);
            if (XHAT) st_out(reinterpret_cast<float4*>(xhat) + o, make_float4(xh[0], xh[1], xh[2], xh[3]));
            carried = r2;
            p1 = q1;
            p2 = q2;
        }
    }
}

template <bool QUANT, bool XHAT>
__global__ void __launch_bounds__(kBThreads, 2)
bn_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ idx,
                   float* __restrict__ xhat, const PoolGeom G, const BnParams2 P,
                   const float* __restrict__ fq_lo, const float* __restrict__ fq_hi, int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    QParams qp = given_qparams(1.0f, 0.0f, 1);
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
        // the integer code orders like the dequantised value iff the table is strictly increasing (it is
        // whenever zp is small, i.e. always for post-ReLU ranges; a huge |lo|/range ratio can merge entries)
        bool inc = true;
        for (int j = threadIdx.x; j + 1 < (1 << fq_k); j += kBThreads) inc = inc && (lut[j] < lut[j + 1]);
        if (__syncthreads_and(inc)) pool_fwd_body<QUANT, true, XHAT>(x, out, idx, xhat, G, P, qp, lut, qh, qmask);
        else pool_fwd_body<QUANT, false, XHAT>(x, out, idx, xhat, G, P, qp, lut, qh, qmask);
    } else {
        pool_fwd_body<false, false, XHAT>(x, out, idx, xhat, G, P, qp, lut, qh, qmask);
    }
}

// A thread owns one 2x2 block of input pixels (rows 2m, 2m+1; columns 2n, 2n+1).  The only windows that cover it
// are (m,n), (m,n+1), (m+1,n), (m+1,n+1): four loads of (grad, argmax code) serve four stores, and the window
// whose centre the block holds -- (m,n), centre (2m,2n) -- is this thread's share of dW / dB.
template <bool REDUCE>
__global__ void __launch_bounds__(kBThreads)
bn_pool_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ gout2, const uint8_t* __restrict__ idx,
                   const float* __restrict__ xhat, float* __restrict__ gx, const PoolGeom G, const BnParams2 P,
                   double* __restrict__ part) {
    __shared__ float red[REDUCE ? 2 * kBThreads * 4 : 1];
    const bool active = (int)threadIdx.x < G.lanes_r * G.cols;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float b, inv; affine2(P, 4 * col + j, a[j], b, inv); }
    }
    const long long blocks = (long long)G.N * G.Ho * G.Wo;
    const float4* g4 = reinterpret_cast<const float4*>(gout);
    const float4* h4 = reinterpret_cast<const float4*>(gout2);      // nullable: second gradient w.r.t. the output
    const uchar4* i4 = reinterpret_cast<const uchar4*>(idx);
    float4* gx4 = reinterpret_cast<float4*>(gx);
    if (active) {
        for (long long o = (long long)blockIdx.x * G.lanes_r + rsub; o < blocks; o += (long long)gridDim.x * G.lanes_r) {
            const int n2 = (int)(o % G.Wo);
            const int m = (int)((o / G.Wo) % G.Ho);
            const long long n = o / ((long long)G.Wo * G.Ho);
            const bool right = n2 + 1 < G.Wo, below = m + 1 < G.Ho;
            // windows in the order a pixel's contributions are summed: (m,n) (m,n+1) (m+1,n) (m+1,n+1)
            const long long base = o * G.cols + col;
            const long long off[4] = {0, (long long)G.cols, (long long)G.Wo * G.cols, ((long long)G.Wo + 1) * G.cols};
            const bool have[4] = {true, right, below, right && below};
            float gs[4][4];
            unsigned char cs[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                uchar4 c = make_uchar4(0, 0, 0, 0);
                if (have[k]) {
                    g = __ldg(g4 + base + off[k]);
                    c = __ldg(i4 + base + off[k]);
                    if (h4) {             // the pooled tensor fed two consumers: sum their gradients here
                        const float4 t = __ldg(h4 + base + off[k]);
                        g.x = __fadd_rn(g.x, t.x); g.y = __fadd_rn(g.y, t.y); g.z = __fadd_rn(g.z, t.z); g.w = __fadd_rn(g.w, t.w);
                    }
                }
                gs[k][0] = g.x; gs[k][1] = g.y; gs[k][2] = g.z; gs[k][3] = g.w;
                cs[k][0] = c.x; cs[k][1] = c.y; cs[k][2] = c.z; cs[k][3] = c.w;
            }
            float xh[4] = {0.f, 0.f, 0.f, 0.f};
            if (REDUCE) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(xhat) + base);
                xh[0] = t.x; xh[1] = t.y; xh[2] = t.z; xh[3] = t.w;
            }
            float p00[4], p01[4], p10[4], p11[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // a code with the ReLU bit clear (< 128) never matches: the gradient stops at an inactive ReLU
                // window-local index of each pixel of the block inside each window:
                //   (2m,2n): centre 4 of W0      (2m,2n+1): 5 of W0, 3 of W1
                //   (2m+1,2n): 7 of W0, 1 of W2  (2m+1,2n+1): 8 of W0, 6 of W1, 2 of W2, 0 of W3
                const int c0 = cs[0][j], c1 = cs[1][j], c2 = cs[2][j], c3 = cs[3][j];
                p00[j] = (c0 == (128 | 4)) ? gs[0][j] : 0.f;
                float t = 0.f;
                if (c0 == (128 | 5)) t += gs[0][j];
                if (c1 == (128 | 3)) t += gs[1][j];
                p01[j] = t;
                t = 0.f;
                if (c0 == (128 | 7)) t += gs[0][j];
                if (c2 == (128 | 1)) t += gs[2][j];
                p10[j] = t;
                t = 0.f;
                if (c0 == (128 | 8)) t += gs[0][j];
                if (c1 == (128 | 6)) t += gs[1][j];
                if (c2 == (128 | 2)) t += gs[2][j];
                if (c3 == (128 | 0)) t += gs[3][j];
                p11[j] = t;
                if (REDUCE && (c0 & 128)) { sb[j] += gs[0][j]; sw[j] = fmaf(gs[0][j], xh[j], sw[j]); }
            }
            const int h = 2 * m, w = 2 * n2;
            float4* dst = gx4 + ((n * G.H + h) * G.W + w) * G.cols + col;
            st_out(dst, make_float4(p00[0] * a[0], p00[1] * a[1], p00[2] * a[2], p00[3] * a[3]));
            if (w + 1 < G.W) st_out(dst + G.cols, make_float4(p01[0] * a[0], p01[1] * a[1], p01[2] * a[2], p01[3] * a[3]));
            if (h + 1 < G.H) {
                dst += (long long)G.W * G.cols;
                st_out(dst, make_float4(p10[0] * a[0], p10[1] * a[1], p10[2] * a[2], p10[3] * a[3]));
                if (w + 1 < G.W) st_out(dst + G.cols, make_float4(p11[0] * a[0], p11[1] * a[1], p11[2] * a[2], p11[3] * a[3]));
            }
        }
    }
    if (REDUCE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[threadIdx.x * 4 + j] = active ? sb[j] : 0.f;
            red[kBThreads * 4 + threadIdx.x * 4 + j] = active ? sw[j] : 0.f;
        }
        __syncthreads();
        if (active && rsub == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float tb = 0.f, tw = 0.f;
                for (int l = 0; l < G.lanes_r; ++l) {
                    tb += red[(l * G.cols + col) * 4 + j];
                    tw += red[kBThreads * 4 + (l * G.cols + col) * 4 + j];
                }
                double* q = part + ((size_t)blockIdx.x * G.C + 4 * col + j) * 2;
                q[0] = (double)tw;      // dW: xhat already carries 1/sqrt(var+eps)
                q[1] = (double)tb;      // dB
            }
        }
    }
}

// ---- TMA-staged backward (default when a window row of every operand is a 16-byte multiple) ---------------
// The gather kernel above loads every (grad, argmax code) element four times (once per 2x2 block that a window
// touches): 4x the LSU instructions and L1/L2 traffic of the 1/4-size operands, 66-76 % of the HBM rate.  Here a
// persistent CTA (one per SM) walks segments of consecutive window rows through a ring of shared-memory slots: one
// thread issues `cp.async.bulk` copies of the NEXT window rows (grad, second grad, codes, x-hat: contiguous runs of
// Wo*C elements) while all threads expand the current row pair from shared memory into two full-resolution rows of
// grad_x (direct 128-bit streaming stores).  Every operand element crosses L2->SM once (plus one halo row per
// segment).  Arithmetic and summation order are the gather kernel's, so results are bit-identical.
constexpr int kPbSlots = 4;
constexpr int kPbMaxThreads = 512;

struct PoolBwdPlan {
    int seg, nseg;            // window rows per segment, segments per image
    int lanes;                // window columns processed per pass (blockDim = lanes * cols)
    int row_f, row_b;         // floats / bytes of one window row of a float operand / of the codes
    int slot_bytes;           // one ring slot: grad [+ grad2] + codes [+ xhat]
    int off_g2, off_idx, off_xh;
};

template <bool REDUCE, bool TWO>
__global__ void __launch_bounds__(kPbMaxThreads, 1)
bn_pool_bwd_tma_kernel(const float* __restrict__ gout, const float* __restrict__ gout2, const uint8_t* __restrict__ idx,
                       const float* __restrict__ xhat, float* __restrict__ gx, const PoolGeom G, const PoolBwdPlan L,
                       const BnParams2 P, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char pb_smem[];
    __shared__ __align__(8) uint64_t full[kPbSlots];
    __shared__ float red[REDUCE ? 2 * kPbMaxThreads * 4 : 1];
    const int col = threadIdx.x % G.cols, lane_w = threadIdx.x / G.cols;
    float a[4], sb[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { float b, inv; affine2(P, 4 * col + j, a[j], b, inv); }

    // this CTA's items t = 0, 1, ...: item = (image n, segment sg); its loads are window rows m0 .. m_last (the last
    // one a halo row when the segment is not the image's last), its steps are window rows m0 .. m1 - 1
    const long long items = (long long)G.N * L.nseg;
    const long long mine = items > blockIdx.x ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto item_rows = [&](long long t, long long& n, int& m0, int& steps, int& loads) {
        const long long it = blockIdx.x + t * gridDim.x;
        n = it / L.nseg;
        m0 = (int)(it % L.nseg) * L.seg;
        const int m1 = min(G.Ho, m0 + L.seg);
        steps = m1 - m0;
        loads = steps + (m1 < G.Ho ? 1 : 0);
    };
    const uint32_t fbytes = (uint32_t)L.row_f * 4u;
    auto issue = [&](long long q, long long n, int m) {             // thread 0: window row m of image n into slot q % K
        const int sl = (int)(q % kPbSlots);
        unsigned char* s = pb_smem + (size_t)sl * L.slot_bytes;
        const long long row = (n * G.Ho + m) * (long long)L.row_f;  // element offset of the row in every operand
        mbar_arrive_expect_tx(&full[sl], fbytes * (1u + (TWO ? 1u : 0u) + (REDUCE ? 1u : 0u)) + (uint32_t)L.row_b);
        bulk_g2s(s, gout + row, fbytes, &full[sl]);
        if (TWO) bulk_g2s(s + L.off_g2, gout2 + row, fbytes, &full[sl]);
        bulk_g2s(s + L.off_idx, idx + row, (uint32_t)L.row_b, &full[sl]);
        if (REDUCE) bulk_g2s(s + L.off_xh, xhat + row, fbytes, &full[sl]);
    };

    if (threadIdx.x == 0) {
        for (int k = 0; k < kPbSlots; ++k) mbar_init(&full[k], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // producer cursor (thread 0 only uses it): next load = row `pm` of item `pt`, global load index `pq`
    long long pt = 0, pq = 0, pn = 0;
    int pm = 0, pleft = 0, psteps = 0;
    if (mine > 0) { int m0; item_rows(0, pn, m0, psteps, pleft); pm = m0; }
    auto produce_upto = [&](long long limit) {                      // issue loads while their index is < limit
        while (pt < mine && pq < limit) {
            issue(pq, pn, pm);
            ++pq; ++pm; --pleft;
            if (pleft == 0) {
                ++pt;
                if (pt < mine) { int m0; item_rows(pt, pn, m0, psteps, pleft); pm = m0; }
            }
        }
    };
    if (threadIdx.x == 0) produce_upto(kPbSlots);

    long long q = 0;                                                 // load index of the current step's own row
    for (long long t = 0; t < mine; ++t) {
        long long n; int m0, steps, loads;
        item_rows(t, n, m0, steps, loads);
        for (int j = 0; j < steps; ++j, ++q) {
            const int m = m0 + j;
            const bool below = m + 1 < G.Ho;                          // then row m+1 is load q+1 of this item
            const int s0 = (int)(q % kPbSlots), s1 = (int)((q + 1) % kPbSlots);
            mbar_wait(&full[s0], (uint32_t)((q / kPbSlots) & 1));
            if (below) mbar_wait(&full[s1], (uint32_t)(((q + 1) / kPbSlots) & 1));
            const unsigned char* r0 = pb_smem + (size_t)s0 * L.slot_bytes;
            const unsigned char* r1 = pb_smem + (size_t)s1 * L.slot_bytes;
            const int h = 2 * m;
            for (int wo = lane_w; wo < G.Wo; wo += L.lanes) {
                const bool right = wo + 1 < G.Wo;
                // windows in the order a pixel's contributions are summed: (m,wo) (m,wo+1) (m+1,wo) (m+1,wo+1)
                const unsigned char* rows[4] = {r0, r0, r1, r1};
                const int wcol[4] = {wo, wo + 1, wo, wo + 1};
                const bool have[4] = {true, right, below, right && below};
                float gs[4][4];
                unsigned char cs[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    uchar4 c = make_uchar4(0, 0, 0, 0);
                    if (have[k]) {
                        const int e = wcol[k] * G.cols + col;
                        g = reinterpret_cast<const float4*>(rows[k])[e];
                        c = reinterpret_cast<const uchar4*>(rows[k] + L.off_idx)[e];
                        if (TWO) {
                            const float4 u = reinterpret_cast<const float4*>(rows[k] + L.off_g2)[e];
                            g.x = __fadd_rn(g.x, u.x); g.y = __fadd_rn(g.y, u.y); g.z = __fadd_rn(g.z, u.z); g.w = __fadd_rn(g.w, u.w);
                        }
                    }
                    gs[k][0] = g.x; gs[k][1] = g.y; gs[k][2] = g.z; gs[k][3] = g.w;
                    cs[k][0] = c.x; cs[k][1] = c.y; cs[k][2] = c.z; cs[k][3] = c.w;
                }
                float xh[4] = {0.f, 0.f, 0.f, 0.f};
                if (REDUCE) {
                    const float4 u = reinterpret_cast<const float4*>(r0 + L.off_xh)[wo * G.cols + col];
                    xh[0] = u.x; xh[1] = u.y; xh[2] = u.z; xh[3] = u.w;
                }
                float p00[4], p01[4], p10[4], p11[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int c0 = cs[0][jj], c1 = cs[1][jj], c2 = cs[2][jj], c3 = cs[3][jj];
                    p00[jj] = (c0 == (128 | 4)) ? gs[0][jj] : 0.f;
                    float v = 0.f;
                    if (c0 == (128 | 5)) v += gs[0][jj];
                    if (c1 == (128 | 3)) v += gs[1][jj];
                    p01[jj] = v;
                    v = 0.f;
                    if (c0 == (128 | 7)) v += gs[0][jj];
                    if (c2 == (128 | 1)) v += gs[2][jj];
                    p10[jj] = v;
                    v = 0.f;
                    if (c0 == (128 | 8)) v += gs[0][jj];
                    if (c1 == (128 | 6)) v += gs[1][jj];
                    if (c2 == (128 | 2)) v += gs[2][jj];
                    if (c3 == (128 | 0)) v += gs[3][jj];
                    p11[jj] = v;
                    if (REDUCE && (c0 & 128)) { sb[jj] += gs[0][jj]; sw[jj] = fmaf(gs[0][jj], xh[jj], sw[jj]); }
                }
                const int w = 2 * wo;
                float4* dst = reinterpret_cast<float4*>(gx) + ((n * G.H + h) * G.W + w) * G.cols + col;
                st_out(dst, make_float4(p00[0] * a[0], p00[1] * a[1], p00[2] * a[2], p00[3] * a[3]));
                if (w + 1 < G.W) st_out(dst + G.cols, make_float4(p01[0] * a[0], p01[1] * a[1], p01[2] * a[2], p01[3] * a[3]));
                if (h + 1 < G.H) {
                    dst += (long long)G.W * G.cols;
                    st_out(dst, make_float4(p10[0] * a[0], p10[1] * a[1], p10[2] * a[2], p10[3] * a[3]));
                    if (w + 1 < G.W) st_out(dst + G.cols, make_float4(p11[0] * a[0], p11[1] * a[1], p11[2] * a[2], p11[3] * a[3]));
                }
            }
            __syncthreads();                                         // every thread is done with row q (and, at an item's
            // last step, with its halo row q+1): those slots may be refilled
            const bool item_end = j + 1 == steps;
            if (threadIdx.x == 0) produce_upto(q + 1 + (item_end && loads > steps ? 1 : 0) + kPbSlots);
        }
        if (loads > steps) ++q;                                      // the halo row was a load, not a step
    }
    if (REDUCE) {
        const int nth = L.lanes * G.cols;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[threadIdx.x * 4 + j] = sb[j];
            red[kPbMaxThreads * 4 + threadIdx.x * 4 + j] = sw[j];
        }
        __syncthreads();
        if (lane_w == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float tb = 0.f, tw = 0.f;
                for (int t2 = col; t2 < nth; t2 += G.cols) {
                    tb += red[t2 * 4 + j];
                    tw += red[kPbMaxThreads * 4 + t2 * 4 + j];
                }
                double* qd = part + ((size_t)blockIdx.x * G.C + 4 * col + j) * 2;
                qd[0] = (double)tw;      // dW: xhat already carries 1/sqrt(var+eps)
                qd[1] = (double)tb;      // dB
            }
        }
    }
}

// plan of the TMA-staged backward, or false when the geometry does not fit (the gather kernel takes over)
static bool make_pool_bwd_plan(const PoolGeom& G, bool two, bool reduce, PoolBwdPlan& L, size_t& smem, int& threads) {
    const long long row_f = (long long)G.Wo * G.C;
    if (G.cols > kPbMaxThreads || (row_f & 15) || row_f > (1 << 20)) return false;       // code row: 16-byte multiple
    L.row_f = (int)row_f;
    L.row_b = (int)row_f;
    int off = L.row_f * 4;
    L.off_g2 = off; if (two) off += L.row_f * 4;
    L.off_idx = off; off += L.row_b;
    L.off_xh = off; if (reduce) off += L.row_f * 4;
    L.slot_bytes = (off + 127) / 128 * 128;
    smem = (size_t)kPbSlots * L.slot_bytes;
    if (smem > 200 * 1024) return false;
    const int max_lanes = kPbMaxThreads / G.cols;
    const int passes = (G.Wo + max_lanes - 1) / max_lanes;
    L.lanes = (G.Wo + passes - 1) / passes;
    threads = L.lanes * G.cols;
    // segments per image: the split whose last round of CTAs is fullest, counting the halo row each segment re-reads
    double best = -1.0;
    L.nseg = 1;
    for (int ns = 1; ns <= 16 && ns <= G.Ho; ++ns) {
        const int seg = (G.Ho + ns - 1) / ns;
        const int real = (G.Ho + seg - 1) / seg;
        const double items = (double)G.N * real;
        const double rounds = (double)(long long)((items + kNumSM - 1) / kNumSM);
        const double eff = items / (rounds * kNumSM) * ((double)seg / (seg + 0.4));
        if (eff > best + 1e-9) { best = eff; L.nseg = real; L.seg = seg; }
    }
    return true;
}

static int make_pool_geom(int N, int C, int H, int W, PoolGeom& G) {
    if (C % 4 != 0 || C / 4 > kBThreads) return OODFQ_EINVAL;
    G.N = N; G.C = C; G.H = H; G.W = W;
    G.Ho = (H - 1) / 2 + 1;
    G.Wo = (W - 1) / 2 + 1;
    G.cols = C / 4;
    G.lanes_r = kBThreads / G.cols;
    G.seg = G.Ho;
    G.nseg = 1;
    return OODFQ_OK;
}

}  // namespace oodfq

#include "bn_pool_ring.cuh"

using namespace oodfq;

extern "C" int oodfq_bn_pool_forward(const float* x, float* out, uint8_t* idx, float* xhat, int N, int C, int H,
                                     int W, const float* weight, const float* bias, const float* running_mean,
                                     const float* running_var, float eps, int flags, const float* fq_lo,
                                     const float* fq_hi, int fq_k, oodfq_stream_t stream) {
    if (!x || !out || !idx || !running_mean || !running_var) return fail(OODFQ_EINVAL, "bn_pool_forward: null pointer");
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(OODFQ_EINVAL, "bn_pool_forward: empty tensor");
    if (!(flags & OODFQ_BN_NHWC) || !(flags & OODFQ_BN_RELU))
        return fail(OODFQ_EINVAL, "bn_pool_forward: only the channels_last BN -> ReLU -> [QuantAct] -> MaxPool(3,2,1) stem is implemented");
    const bool quant = flags & OODFQ_BN_QUANT;
    if (quant && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8)) return fail(OODFQ_EINVAL, "bn_pool_forward: fake-quant needs a range and k in [1,8]");
    PoolGeom G;
    if (make_pool_geom(N, C, H, W, G) != OODFQ_OK || !aligned16(x) || !aligned16(out) || (xhat && !aligned16(xhat)) ||
        (reinterpret_cast<uintptr_t>(idx) & 3u))
        return fail(OODFQ_EINVAL, "bn_pool_forward: needs C %% 4 == 0, C <= 1024 and aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const BnParams2 P{weight, bias, running_mean, running_var, eps};
    PoolFwdPlan L;
    size_t smem = 0;
    int threads = 0;
    if (!(flags & OODFQ_BN_POOL_REGISTER) && make_pool_fwd_plan(G, L, smem, threads)) {
        const void* kernels[4] = {(const void*)bn_pool_fwd_tma_kernel<false, false>, (const void*)bn_pool_fwd_tma_kernel<false, true>,
                                  (const void*)bn_pool_fwd_tma_kernel<true, false>, (const void*)bn_pool_fwd_tma_kernel<true, true>};
        const int v = (quant ? 2 : 0) + (xhat ? 1 : 0);
        static size_t smem_set[4] = {0, 0, 0, 0};
        bool ok = true;
        if (smem > smem_set[v]) {
            ok = cudaFuncSetAttribute(kernels[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
            if (ok) smem_set[v] = smem; else (void)cudaGetLastError();
        }
        if (ok) {
            const long long items = (long long)N * L.nseg;
            const unsigned grid = (unsigned)(items < kNumSM ? items : kNumSM);
            if (v == 0) bn_pool_fwd_tma_kernel<false, false><<<grid, threads, smem, st>>>(x, out, idx, xhat, G, L, P, fq_lo, fq_hi, fq_k);
            else if (v == 1) bn_pool_fwd_tma_kernel<false, true><<<grid, threads, smem, st>>>(x, out, idx, xhat, G, L, P, fq_lo, fq_hi, fq_k);
            else if (v == 2) bn_pool_fwd_tma_kernel<true, false><<<grid, threads, smem, st>>>(x, out, idx, xhat, G, L, P, fq_lo, fq_hi, fq_k);
            else bn_pool_fwd_tma_kernel<true, true><<<grid, threads, smem, st>>>(x, out, idx, xhat, G, L, P, fq_lo, fq_hi, fq_k);
            count_launch();
            return check_launch("bn_pool_forward(tma)");
        }
    }
    static const int occ[4] = {resident_ctas(bn_pool_fwd_kernel<false, false>, kBThreads), resident_ctas(bn_pool_fwd_kernel<false, true>, kBThreads),
                               resident_ctas(bn_pool_fwd_kernel<true, false>, kBThreads), resident_ctas(bn_pool_fwd_kernel<true, true>, kBThreads)};
    const int per_sm = occ[(quant ? 2 : 0) + (xhat ? 1 : 0)];
    // work item = lanes_r window columns x one segment of output rows.  Each segment re-reads one input row, so
    // segments stay >= 8 rows; within that, aim for ~6 items per resident CTA so the last wave is short.
    const long long cap = (long long)kNumSM * per_sm;
    const long long groups = ((long long)N * G.Wo + G.lanes_r - 1) / G.lanes_r;
    long long nseg = (6 * cap + groups - 1) / groups;
    const long long max_seg = (G.Ho + 7) / 8;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    G.seg = (int)((G.Ho + nseg - 1) / nseg);
    G.nseg = (G.Ho + G.seg - 1) / G.seg;
    const long long want = groups * G.nseg;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (quant && xhat) bn_pool_fwd_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else if (quant) bn_pool_fwd_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else if (xhat) bn_pool_fwd_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else bn_pool_fwd_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    count_launch();
    return check_launch("bn_pool_forward");
}

extern "C" int oodfq_bn_pool_backward(const float* grad_out, const float* grad_out2, const uint8_t* idx, const float* xhat, float* grad_x,
                                      int N, int C, int H, int W, const float* weight, const float* bias,
                                      const float* running_mean, const float* running_var, float eps,
                                      float* dwdb, void* workspace, oodfq_stream_t stream) {
    if (!grad_out || !idx || !grad_x || !running_mean || !running_var) return fail(OODFQ_EINVAL, "bn_pool_backward: null pointer");
    if (dwdb && (!xhat || !workspace)) return fail(OODFQ_EINVAL, "bn_pool_backward: parameter gradients need xhat and the workspace");
    PoolGeom G;
    if (make_pool_geom(N, C, H, W, G) != OODFQ_OK || !aligned16(grad_out) || !aligned16(grad_x) || (xhat && !aligned16(xhat)) ||
        (grad_out2 && !aligned16(grad_out2)))
        return fail(OODFQ_EINVAL, "bn_pool_backward: needs C %% 4 == 0, C <= 1024 and aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const BnParams2 P{weight, bias, running_mean, running_var, eps};
    PoolBwdPlan L;
    size_t smem = 0;
    int threads = 0;
    if (aligned16(idx) && make_pool_bwd_plan(G, grad_out2 != nullptr, dwdb != nullptr, L, smem, threads)) {
        const void* kernels[4] = {(const void*)bn_pool_bwd_tma_kernel<false, false>, (const void*)bn_pool_bwd_tma_kernel<false, true>,
                                  (const void*)bn_pool_bwd_tma_kernel<true, false>, (const void*)bn_pool_bwd_tma_kernel<true, true>};
        const int v = (dwdb ? 2 : 0) + (grad_out2 ? 1 : 0);
        static size_t smem_set[4] = {0, 0, 0, 0};
        bool ok = true;
        if (smem > smem_set[v]) {
            ok = cudaFuncSetAttribute(kernels[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
            if (ok) smem_set[v] = smem; else (void)cudaGetLastError();
        }
        if (ok) {
            const long long items = (long long)N * L.nseg;
            const unsigned grid = (unsigned)(items < kNumSM ? items : kNumSM);
            double* part = dwdb ? fold_target(ws->bn_partial, C, (int)grid) : nullptr;   // workspace, or its own region (fold.cu)
            if (v == 0) bn_pool_bwd_tma_kernel<false, false><<<grid, threads, smem, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, L, P, part);
            else if (v == 1) bn_pool_bwd_tma_kernel<false, true><<<grid, threads, smem, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, L, P, part);
            else if (v == 2) bn_pool_bwd_tma_kernel<true, false><<<grid, threads, smem, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, L, P, part);
            else bn_pool_bwd_tma_kernel<true, true><<<grid, threads, smem, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, L, P, part);
            count_launch();
            int rc = check_launch("bn_pool_backward(tma)");
            if (rc != OODFQ_OK || !dwdb) return rc;
            return fold_finish(part, ws->bn_partial, C, (int)grid, dwdb, st);
        }
    }
    static const int occ[2] = {resident_ctas(bn_pool_bwd_kernel<false>, kBThreads), resident_ctas(bn_pool_bwd_kernel<true>, kBThreads)};
    const long long pixels = (long long)N * H * W;
    long long want = (pixels + G.lanes_r - 1) / G.lanes_r, cap = (long long)kNumSM * occ[dwdb ? 1 : 0];
    const long long table = (long long)kMaxBnSplit * kMaxBnChannels / C;
    if (dwdb && cap > table) cap = table;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    double* part = dwdb ? fold_target(ws->bn_partial, C, (int)grid) : nullptr;
    if (dwdb) bn_pool_bwd_kernel<true><<<grid, kBThreads, 0, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, P, part);
    else bn_pool_bwd_kernel<false><<<grid, kBThreads, 0, st>>>(grad_out, grad_out2, idx, xhat, grad_x, G, P, part);
    count_launch();
    int rc = check_launch("bn_pool_backward");
    if (rc != OODFQ_OK || !dwdb) return rc;
    return fold_finish(part, ws->bn_partial, C, (int)grid, dwdb, st);
}
