// Eval-mode BatchNorm fused with the ReLU and QuantAct that follow it.
//
// In the reference the student always runs in eval() (trainer_direct.py:411), so every BatchNorm is
// the per-channel affine  z = a_c*x + b_c,  a_c = w_c / sqrt(rv_c + eps),  b_c = bias_c - rm_c*a_c,
// and the `nn.Sequential(ReLU, QuantAct)` that quantize_model puts behind it (main_direct.py:464-465)
// re-reads and re-writes the same tensor twice more.  Measured on the ImageNet step, ATen's eval-mode
// batch_norm_backward_kernel alone is 24 % of the step and cuDNN's inference BN another 5 %
// (profiles/r1_step_share.txt).  Here:
//
//   forward   y = fakequant(relu(a_c*x + b_c))        one read, one write        (8 B/elem)
//   backward  g' = g * [a_c*x + b_c > 0];  grad_x = g' * a_c;                    (12 B/elem)
//             dW_c = sum g' * (x - rm_c) / sqrt(rv_c + eps);  dB_c = sum g'      (same read)
//
// RELU and QUANT are flags: without them this is a plain (and fast) eval-mode BN for the layers that
// feed a residual add.  The fake-quant arithmetic is the reference's (common.cuh), applied to the
// fp32 value z the affine produces; the QuantAct backward is the identity STE.
//
// Decomposition, determinism and grid sizing: bn_geom.cuh.  Roofline: HBM.
#include "bn_geom.cuh"

namespace oodfq {

// =============================================================================== forward
template <bool RELU, bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_plane_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ zdbg, int N, int C,
                    long long HW, int split, const BnParams P, const float* __restrict__ fq_lo,
                    const float* __restrict__ fq_hi, int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    const int c = blockIdx.x / split, sp = blockIdx.x % split;
    QParams qp;
    float lowc = 0.0f;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        lowc = relu_lower_bound(qp);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    float a, b, invstd;
    affine_of(P, c, a, b, invstd);
    const int n4 = (int)(HW >> 2);
    for (int n = sp; n < N; n += split) {
        const long long base = ((long long)n * C + c) * HW;
        const float4* p = reinterpret_cast<const float4*>(x + base);
        float4* q = reinterpret_cast<float4*>(y + base);
        float4* zq = zdbg ? reinterpret_cast<float4*>(zdbg + base) : nullptr;
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * kBThreads) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) v[u] = ld_stream(p + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) {
                    float4 r, z;
                    r.x = head<RELU, QUANT>(v[u].x, a, b, qp, lowc, lut, qh, qmask, z.x);
                    r.y = head<RELU, QUANT>(v[u].y, a, b, qp, lowc, lut, qh, qmask, z.y);
                    r.z = head<RELU, QUANT>(v[u].z, a, b, qp, lowc, lut, qh, qmask, z.z);
                    r.w = head<RELU, QUANT>(v[u].w, a, b, qp, lowc, lut, qh, qmask, z.w);
                    st_out(q + i, r);
                    if (zq) zq[i] = RELU ? make_float4(relu_keep_nan(z.x), relu_keep_nan(z.y), relu_keep_nan(z.z), relu_keep_nan(z.w)) : z;
                }
            }
        }
    }
}

template <int VEC, bool RELU, bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_group_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ zdbg, const BnGeom G,
                    const BnParams P, const float* __restrict__ fq_lo, const float* __restrict__ fq_hi, int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;
    const int hw = (int)((G.cg > 1) ? G.HW : 0x7fffffff);
    QParams qp;
    float lowc = 0.0f;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        lowc = relu_lower_bound(qp);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    const int e0 = threadIdx.x * VEC;
    if (e0 >= len) return;
    float a[VEC], b[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        float invstd;
        affine_of(P, c0 + (e0 + j) / hw, a[j], b[j], invstd);
    }
    for (int n = sp; n < G.N; n += kDepth * G.split) {
        float v[kDepth][VEC];
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (nn < G.N) load_vec<VEC>(x + nn * row + off + e0, v[d]);
        }
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (nn < G.N) {
                float r[VEC], z[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) r[j] = head<RELU, QUANT>(v[d][j], a[j], b[j], qp, lowc, lut, qh, qmask, z[j]);
                store_vec<VEC>(y + nn * row + off + e0, r);
                if (zdbg) {
                    if (RELU) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) z[j] = relu_keep_nan(z[j]);
                    }
                    store_vec<VEC>(zdbg + nn * row + off + e0, z);
                }
            }
        }
    }
}

// =============================================================================== backward
// REDUCE: also accumulate dB_c = sum g', dW_c = sum g' (x - rm_c) * invstd_c as fp64 into dwdb[2C]
template <bool RELU, bool REDUCE>
__global__ void __launch_bounds__(kBThreads)
bn_plane_bwdx_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx, int N,
                     int C, long long HW, int split, const BnParams P, float* __restrict__ dwdb, Workspace* ws) {
    __shared__ float r1[kBThreads / 32], r2[kBThreads / 32];
    __shared__ int s_last;
    const int c = blockIdx.x / split, sp = blockIdx.x % split;
    float a, b, invstd;
    affine_of(P, c, a, b, invstd);
    const float rm = __ldg(P.rm + c);
    const int n4 = (int)(HW >> 2);
    float sb[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n = sp; n < N; n += split) {
        const long long base = ((long long)n * C + c) * HW;
        const float4* p = reinterpret_cast<const float4*>(x + base);
        const float4* pg = reinterpret_cast<const float4*>(gy + base);
        float4* q = reinterpret_cast<float4*>(gx + base);
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * kBThreads) {
            float4 v[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) { v[u] = ld_stream(p + i); g[u] = ld_stream(pg + i); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) {
                    const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                    float gs[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (RELU && !(fmaf(xs[j], a, b) > 0.0f)) gs[j] = 0.0f;
                        if (REDUCE) { sb[j] += gs[j]; sw[j] = fmaf(gs[j], xs[j] - rm, sw[j]); }
                    }
                    st_out(q + i, make_float4(gs[0] * a, gs[1] * a, gs[2] * a, gs[3] * a));
                }
            }
        }
    }
    if (!REDUCE) return;
    float t1 = warp_sum((sb[0] + sb[1]) + (sb[2] + sb[3]));
    float t2 = warp_sum((sw[0] + sw[1]) + (sw[2] + sw[3]));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { r1[warp] = t1; r2[warp] = t2; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kBThreads / 32; ++w) { t1 += r1[w]; t2 += r2[w]; }
        double* pp = ws->bn_partial + ((size_t)sp * C + c) * 2;
        pp[0] = (double)t2 * (double)invstd;   // dW
        pp[1] = (double)t1;                    // dB
        __threadfence();
        s_last = (atomicAdd(&ws->bn_ticket[c], 1) == split - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) fold_partials(ws->bn_partial, C, c, split, lane, dwdb);
    if (threadIdx.x == 0) ws->bn_ticket[c] = 0;
}

template <int VEC, bool RELU, bool REDUCE>
__global__ void __launch_bounds__(kBThreads)
bn_group_bwdx_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                     const BnGeom G, const BnParams P, float* __restrict__ dwdb, Workspace* ws) {
    __shared__ float s1[REDUCE ? kBThreads * VEC : 1];
    __shared__ float s2[REDUCE ? kBThreads * VEC : 1];
    __shared__ int s_last;
    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;
    const int hw = (int)((G.cg > 1) ? G.HW : 0x7fffffff);
    const int e0 = threadIdx.x * VEC;
    const bool active = e0 < len;
    float a[VEC], b[VEC], rm[VEC], sb[VEC], sw[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sb[j] = 0.f; sw[j] = 0.f; a[j] = 0.f; b[j] = 0.f; rm[j] = 0.f;
        if (active) {
            float invstd;
            const int c = c0 + (e0 + j) / hw;
            affine_of(P, c, a[j], b[j], invstd);
            rm[j] = __ldg(P.rm + c);
        }
    }
    if (active) {
        for (int n = sp; n < G.N; n += kDepth * G.split) {
            float v[kDepth][VEC], gg[kDepth][VEC];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const int nn = n + d * G.split;
                if (nn < G.N) {
                    load_vec<VEC>(x + nn * row + off + e0, v[d]);
                    load_vec<VEC>(gy + nn * row + off + e0, gg[d]);
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const int nn = n + d * G.split;
                if (nn < G.N) {
                    float r[VEC];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        float t = gg[d][j];
                        if (RELU && !(fmaf(v[d][j], a[j], b[j]) > 0.0f)) t = 0.0f;
                        if (REDUCE) { sb[j] += t; sw[j] = fmaf(t, v[d][j] - rm[j], sw[j]); }
                        r[j] = t * a[j];
                    }
                    store_vec<VEC>(gx + nn * row + off + e0, r);
                }
            }
        }
    }
    if (!REDUCE) return;
    if (active) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s1[e0 + j] = sb[j]; s2[e0 + j] = sw[j]; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = (G.cg > 1) ? min(G.cg, G.C - c0) : 1;
    const int part = ck * G.split + sp;
    const int nparts = per_group;
    for (int c = warp; c < nch; c += kBThreads / 32) {
        int eb = (G.cg > 1) ? (int)(c * G.HW) : 0;
        int ee = (G.cg > 1) ? (int)((c + 1) * G.HW) : len;
        float t1 = 0.f, t2 = 0.f;
        for (int e = eb + lane; e < ee; e += 32) { t1 += s1[e]; t2 += s2[e]; }
        t1 = warp_sum(t1);
        t2 = warp_sum(t2);
        if (lane == 0) {
            float aa, bb, invstd;
            affine_of(P, c0 + c, aa, bb, invstd);
            double* p = ws->bn_partial + ((size_t)part * G.C + (c0 + c)) * 2;
            p[0] = (double)t2 * (double)invstd;
            p[1] = (double)t1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&ws->bn_ticket[g], 1) == nparts - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = warp; c < nch; c += kBThreads / 32) fold_partials(ws->bn_partial, G.C, c0 + c, nparts, lane, dwdb);
    if (threadIdx.x == 0) ws->bn_ticket[g] = 0;
}

// =============================================================================== NHWC (channels_last)
// x is [R = N*H*W rows][C]; channels are the fastest axis.  cuDNN's tensor-core convolutions are NHWC
// inside: feeding them channels_last tensors removes the nchwToNhwc / nhwcToNchw transposes that are 28 %
// of the fused NCHW step (profiles/r1_step_share_fused.txt).  Mapping: a thread owns one 128-bit column
// (4 consecutive channels, fixed for its lifetime -> coefficients in registers) and walks down the rows;
// the CTA's 256 threads cover `lanes_r = 256 / (C/4)` rows at a time as ONE contiguous run of memory.
template <bool RELU, bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ zdbg,
                   uint8_t* __restrict__ mask, const NhwcGeom G, const BnParams P, const float* __restrict__ fq_lo,
                   const float* __restrict__ fq_hi, int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    // The first rows of this thread are requested BEFORE the parameter prologue (range -> table -> barrier, running
    // statistics -> affine): three dependent memory latencies that a 10 us launch on a small plane otherwise pays
    // in front of its first load (profiles/r2_microbench.txt: 7x7 planes at half the rate of the 56x56 ones).
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    const bool active = (int)threadIdx.x < G.lanes_r * wcols;
    const int rsub = threadIdx.x / wcols;
    const long long rstep = (long long)G.lanes_r * gridDim.x;
    const long long r0 = (long long)blockIdx.x * G.lanes_r + rsub;
    const int col0 = threadIdx.x % wcols;
    float4 v[kDepth];
    if (active) {
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const long long rr = r0 + d * rstep;
            if (rr < G.R) v[d] = ld_stream(reinterpret_cast<const float4*>(x) + rr * G.cols + col0);
        }
    }
    QParams qp;
    float lowc = 0.0f;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        lowc = relu_lower_bound(qp);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    if (!active) return;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + col0;
        if (col >= G.cols) continue;
        float a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { float invstd; affine_of(P, 4 * col + j, a[j], b[j], invstd); }
        for (long long r = r0; r < G.R; r += kDepth * rstep) {
            if (cb != 0 || r != r0) {
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const long long rr = r + d * rstep;
                    if (rr < G.R) v[d] = ld_stream(reinterpret_cast<const float4*>(x) + rr * G.cols + col);
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R) {
                    float4 o, z;
                    o.x = head<RELU, QUANT>(v[d].x, a[0], b[0], qp, lowc, lut, qh, qmask, z.x);
                    o.y = head<RELU, QUANT>(v[d].y, a[1], b[1], qp, lowc, lut, qh, qmask, z.y);
                    o.z = head<RELU, QUANT>(v[d].z, a[2], b[2], qp, lowc, lut, qh, qmask, z.z);
                    o.w = head<RELU, QUANT>(v[d].w, a[3], b[3], qp, lowc, lut, qh, qmask, z.w);
                    st_out(reinterpret_cast<float4*>(y) + rr * G.cols + col, o);
                    if (zdbg)
                        reinterpret_cast<float4*>(zdbg)[rr * G.cols + col] =
                            RELU ? make_float4(relu_keep_nan(z.x), relu_keep_nan(z.y), relu_keep_nan(z.z), relu_keep_nan(z.w)) : z;
                    // one byte per 128-bit column: bit j = the ReLU of channel j is open (what the backward re-derives
                    // from x otherwise: a*x + b > 0)
                    if (RELU && mask)
                        mask[rr * G.cols + col] = (uint8_t)((z.x > 0.0f ? 1 : 0) | (z.y > 0.0f ? 2 : 0) | (z.z > 0.0f ? 4 : 0) |
                                                            (z.w > 0.0f ? 8 : 0));
                }
            }
        }
    }
}

// Backward through BN + ReLU (+ identity STE) when only grad_x is wanted and the forward left its ReLU mask:
// grad_x = [open] * grad_y * a_c -- x is not read at all (8.25 instead of 12 B/elem).
__global__ void __launch_bounds__(kBThreads)
bn_nhwc_bwdx_mask_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ gy, float* __restrict__ gx,
                         const NhwcGeom G, const BnParams P) {
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    if ((int)threadIdx.x >= G.lanes_r * wcols) return;
    const int rsub = threadIdx.x / wcols;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + threadIdx.x % wcols;
        if (col >= G.cols) continue;
        float a[4], b, invstd;
#pragma unroll
        for (int j = 0; j < 4; ++j) affine_of(P, 4 * col + j, a[j], b, invstd);
        const long long rstep = (long long)G.lanes_r * gridDim.x;
        for (long long r = (long long)blockIdx.x * G.lanes_r + rsub; r < G.R; r += kDepth * rstep) {
            float4 g[kDepth];
            unsigned open[kDepth];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R) {
                    g[d] = ld_stream(reinterpret_cast<const float4*>(gy) + rr * G.cols + col);
                    open[d] = __ldg(mask + rr * G.cols + col);
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R)
                    st_out(reinterpret_cast<float4*>(gx) + rr * G.cols + col,
                           make_float4(((open[d] & 1u) ? g[d].x : 0.0f) * a[0], ((open[d] & 2u) ? g[d].y : 0.0f) * a[1],
                                       ((open[d] & 4u) ? g[d].z : 0.0f) * a[2], ((open[d] & 8u) ? g[d].w : 0.0f) * a[3]));
            }
        }
    }
}

// Backward through a BatchNorm whose INPUT is also tapped by the BN-statistics loss (every BatchNorm of the
// distillation loop, distill_data.py:69-78 / :252-265): the chain is  grad_bn = [open] * grad_y * a_c  (kernel above)
// followed by  grad_x = grad_bn + g * (gmean_c / M + gvar_c * 2 (x - mean_c) / M)  (bn_nhwc_bwd_kernel of bn_stats.cu):
// 8-12 + 12 B/elem.  x is read by the second step anyway, so one pass does both -- 12 B/elem -- with the same
// roundings in the same order (bit-identical to the chain).
template <bool RELU>
__global__ void __launch_bounds__(kBThreads)
bn_nhwc_bwdx_tap_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                        const NhwcGeom G, const BnParams P, const float* __restrict__ mean,
                        const float* __restrict__ gmean, const float* __restrict__ gvar, float inv_count,
                        const float* __restrict__ gscale) {
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    if ((int)threadIdx.x >= G.lanes_r * wcols) return;
    const int rsub = threadIdx.x / wcols;
    const float gs = gscale ? __ldg(gscale) : 1.0f;
    const long long rstep = (long long)G.lanes_r * gridDim.x;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + threadIdx.x % wcols;
        if (col >= G.cols) continue;
        float a[4], b[4], ca[4], cbv[4], mu[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float invstd;
            affine_of(P, 4 * col + j, a[j], b[j], invstd);
            ca[j] = gs * 2.0f * __ldg(gvar + 4 * col + j) * inv_count;
            cbv[j] = gs * __ldg(gmean + 4 * col + j) * inv_count;
            mu[j] = __ldg(mean + 4 * col + j);
        }
        for (long long r = (long long)blockIdx.x * G.lanes_r + rsub; r < G.R; r += kDepth * rstep) {
            float4 v[kDepth], g[kDepth];
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R) {
                    v[d] = ld_stream(reinterpret_cast<const float4*>(x) + rr * G.cols + col);
                    g[d] = ld_stream(reinterpret_cast<const float4*>(gy) + rr * G.cols + col);
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R) {
                    const float xs[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
                    float gsv[4] = {g[d].x, g[d].y, g[d].z, g[d].w}, o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (RELU && !(fmaf(xs[j], a[j], b[j]) > 0.0f)) gsv[j] = 0.0f;
                        o[j] = __fadd_rn(fmaf(ca[j], xs[j] - mu[j], cbv[j]), __fmul_rn(gsv[j], a[j]));
                    }
                    st_out(reinterpret_cast<float4*>(gx) + rr * G.cols + col, make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
    }
}

// Backward: every CTA leaves one fp64 partial (dW, dB) per channel in ws->bn_partial[cta][C][2];
// bn_nhwc_fold_kernel then sums the partials of each channel in CTA order (deterministic).
template <bool RELU, bool REDUCE>
__global__ void __launch_bounds__(kBThreads)
bn_nhwc_bwdx_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                    const NhwcGeom G, const BnParams P, double* __restrict__ part) {
    __shared__ __align__(16) float red[REDUCE ? 2 * kBThreads * 4 : 4];
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    const bool active = (int)threadIdx.x < G.lanes_r * wcols;
    const int rsub = threadIdx.x / wcols;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + threadIdx.x % wcols;
        const bool on = active && col < G.cols;
        float a[4], b[4], rm[4], inv[4], sb[4], sw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sb[j] = 0.f; sw[j] = 0.f; a[j] = 0.f; b[j] = 0.f; rm[j] = 0.f; inv[j] = 0.f;
            if (on) { affine_of(P, 4 * col + j, a[j], b[j], inv[j]); rm[j] = __ldg(P.rm + 4 * col + j); }
        }
        if (on) {
            const long long rstep = (long long)G.lanes_r * gridDim.x;
            for (long long r = (long long)blockIdx.x * G.lanes_r + rsub; r < G.R; r += kDepth * rstep) {
                float4 v[kDepth], g[kDepth];
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const long long rr = r + d * rstep;
                    if (rr < G.R) {
                        v[d] = ld_stream(reinterpret_cast<const float4*>(x) + rr * G.cols + col);
                        g[d] = ld_stream(reinterpret_cast<const float4*>(gy) + rr * G.cols + col);
                    }
                }
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const long long rr = r + d * rstep;
                    if (rr < G.R) {
                        const float xs[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
                        float gs[4] = {g[d].x, g[d].y, g[d].z, g[d].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (RELU && !(fmaf(xs[j], a[j], b[j]) > 0.0f)) gs[j] = 0.0f;
                            if (REDUCE) { sb[j] += gs[j]; sw[j] = fmaf(gs[j], xs[j] - rm[j], sw[j]); }
                        }
                        st_out(reinterpret_cast<float4*>(gx) + rr * G.cols + col,
                               make_float4(gs[0] * a[0], gs[1] * a[1], gs[2] * a[2], gs[3] * a[3]));
                    }
                }
            }
        }
        if (REDUCE) {
            // fold the CTA's row-lanes (fixed tree over the lanes), then one fp64 partial per channel
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                red[threadIdx.x * 4 + j] = on ? sb[j] : 0.f;
                red[kBThreads * 4 + threadIdx.x * 4 + j] = on ? sw[j] : 0.f;
            }
            lane_tree_fold<2>(red, rsub, wcols, G.lanes_r);
            if (on && rsub == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double* p = part + ((size_t)blockIdx.x * G.C + 4 * col + j) * 2;
                    p[0] = (double)red[kBThreads * 4 + threadIdx.x * 4 + j] * (double)inv[j];
                    p[1] = (double)red[threadIdx.x * 4 + j];
                }
            }
        }
    }
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_bn_eval_forward(const float* x, float* y, float* z_debug, int N, int C, long long HW,
                                     const float* weight, const float* bias, const float* running_mean,
                                     const float* running_var, float eps, int flags, const float* fq_lo,
                                     const float* fq_hi, int fq_k, uint8_t* relu_mask, oodfq_stream_t stream) {
    if (!x || !y || !running_mean || !running_var) return fail(OODFQ_EINVAL, "bn_eval_forward: null pointer");
    if (relu_mask && !((flags & OODFQ_BN_NHWC) && (flags & OODFQ_BN_RELU)))
        return fail(OODFQ_EINVAL, "bn_eval_forward: the ReLU mask output exists for channels_last tensors with OODFQ_BN_RELU only");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "bn_eval_forward: empty tensor");
    const bool relu = flags & OODFQ_BN_RELU, quant = flags & OODFQ_BN_QUANT;
    if (quant && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8))
        return fail(OODFQ_EINVAL, "bn_eval_forward: fused fake-quant needs a range and k in [1,8]");
    cudaStream_t st = (cudaStream_t)stream;
    const BnParams P{weight, bias, running_mean, running_var, eps};
    const bool vec_ok = aligned16(x) && aligned16(y) && (!z_debug || aligned16(z_debug));
    if (flags & OODFQ_BN_NHWC) {
        if (!vec_ok || (C % 4) != 0) return fail(OODFQ_EINVAL, "bn_eval_forward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        const NhwcGeom G = make_nhwc((long long)N * HW, C);
        static const int per_sm = resident_ctas(bn_nhwc_fwd_kernel<true, true>, kBThreads);
        long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
        long long cap = (long long)kNumSM * per_sm;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        if (relu && quant) bn_nhwc_fwd_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, relu_mask, G, P, fq_lo, fq_hi, fq_k);
        else if (relu) bn_nhwc_fwd_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, relu_mask, G, P, fq_lo, fq_hi, fq_k);
        else if (quant) bn_nhwc_fwd_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, relu_mask, G, P, fq_lo, fq_hi, fq_k);
        else bn_nhwc_fwd_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, relu_mask, G, P, fq_lo, fq_hi, fq_k);
        count_launch();
        return check_launch("bn_eval_forward");
    }
    if (plane_ok(HW, vec_ok)) {
        static const int per_sm = resident_ctas(bn_plane_fwd_kernel<true, true>, kBThreads);
        const int split = pick_split(C, N, 1 << 20, kNumSM * per_sm);
        const unsigned grid = (unsigned)C * split;
        if (relu && quant) bn_plane_fwd_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, N, C, HW, split, P, fq_lo, fq_hi, fq_k);
        else if (relu) bn_plane_fwd_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, N, C, HW, split, P, fq_lo, fq_hi, fq_k);
        else if (quant) bn_plane_fwd_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, N, C, HW, split, P, fq_lo, fq_hi, fq_k);
        else bn_plane_fwd_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, N, C, HW, split, P, fq_lo, fq_hi, fq_k);
    } else {
        BnGeom G; int vec; char why[128];
        static const int per_sm = resident_ctas(bn_group_fwd_kernel<4, true, true>, kBThreads);
        if (make_geom(N, C, HW, vec_ok, kNumSM * per_sm, G, vec, why, sizeof(why)) != OODFQ_OK)
            return fail(OODFQ_EINVAL, "bn_eval_forward: %s", why);
        const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
        if (vec == 4) {
            if (relu && quant) bn_group_fwd_kernel<4, true, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else if (relu) bn_group_fwd_kernel<4, true, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else if (quant) bn_group_fwd_kernel<4, false, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else bn_group_fwd_kernel<4, false, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
        } else {
            if (relu && quant) bn_group_fwd_kernel<1, true, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else if (relu) bn_group_fwd_kernel<1, true, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else if (quant) bn_group_fwd_kernel<1, false, true><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
            else bn_group_fwd_kernel<1, false, false><<<grid, kBThreads, 0, st>>>(x, y, z_debug, G, P, fq_lo, fq_hi, fq_k);
        }
    }
    count_launch();
    return check_launch("bn_eval_forward");
}

extern "C" int oodfq_bn_eval_backward(const float* x, const float* grad_y, float* grad_x, int N, int C,
                                      long long HW, const float* weight, const float* bias,
                                      const float* running_mean, const float* running_var, float eps,
                                      int flags, float* dwdb, void* workspace, const uint8_t* relu_mask,
                                      oodfq_stream_t stream) {
    const bool by_mask = relu_mask && (flags & OODFQ_BN_NHWC) && (flags & OODFQ_BN_RELU) && !dwdb;
    if ((!x && !by_mask) || !grad_y || !grad_x || !running_mean || !running_var) return fail(OODFQ_EINVAL, "bn_eval_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "bn_eval_backward: empty tensor");
    if (dwdb && !workspace) return fail(OODFQ_EINVAL, "bn_eval_backward: parameter gradients need the workspace");
    if (C > kMaxBnChannels) return fail(OODFQ_EINVAL, "bn_eval_backward: C=%d exceeds %d", C, kMaxBnChannels);
    const bool relu = flags & OODFQ_BN_RELU, reduce = dwdb != nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const BnParams P{weight, bias, running_mean, running_var, eps};
    const bool vec_ok = (!x || aligned16(x)) && aligned16(grad_y) && aligned16(grad_x);
    if (flags & OODFQ_BN_NHWC) {
        if (!vec_ok || (C % 4) != 0) return fail(OODFQ_EINVAL, "bn_eval_backward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        const NhwcGeom G = make_nhwc((long long)N * HW, C);
        if (by_mask) {
            static const int per_sm = resident_ctas(bn_nhwc_bwdx_mask_kernel, kBThreads);
            long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
            const long long cap = (long long)kNumSM * per_sm;
            bn_nhwc_bwdx_mask_kernel<<<(unsigned)(want < 1 ? 1 : (want < cap ? want : cap)), kBThreads, 0, st>>>(relu_mask, grad_y, grad_x, G, P);
            count_launch();
            return check_launch("bn_eval_backward(mask)");
        }
        // resident CTAs differ a lot between the variants (97 vs 48 registers): size each grid by its own
        static const int occ[4] = {resident_ctas(bn_nhwc_bwdx_kernel<false, false>, kBThreads),
                                   resident_ctas(bn_nhwc_bwdx_kernel<false, true>, kBThreads),
                                   resident_ctas(bn_nhwc_bwdx_kernel<true, false>, kBThreads),
                                   resident_ctas(bn_nhwc_bwdx_kernel<true, true>, kBThreads)};
        long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
        long long cap = (long long)kNumSM * occ[(relu ? 2 : 0) + (reduce ? 1 : 0)];
        const long long table = (long long)kMaxBnSplit * kMaxBnChannels / C;     // partial slots that fit
        if (reduce && cap > table) cap = table;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        // the per-CTA partials go to the workspace and are folded by a launch right behind this one, or -- while folds
        // are deferred (fold.cu) -- to a region of their own, folded with everybody else's at the flush
        double* part = reduce ? fold_target(ws->bn_partial, C, (int)grid) : nullptr;
        if (relu && reduce) bn_nhwc_bwdx_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, part);
        else if (relu) bn_nhwc_bwdx_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, part);
        else if (reduce) bn_nhwc_bwdx_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, part);
        else bn_nhwc_bwdx_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, part);
        count_launch();
        int rc = check_launch("bn_eval_backward");
        if (rc != OODFQ_OK || !reduce) return rc;
        return fold_finish(part, ws->bn_partial, C, (int)grid, dwdb, st);
    }
    if (plane_ok(HW, vec_ok)) {
        static const int occ[4] = {resident_ctas(bn_plane_bwdx_kernel<false, false>, kBThreads),
                                   resident_ctas(bn_plane_bwdx_kernel<false, true>, kBThreads),
                                   resident_ctas(bn_plane_bwdx_kernel<true, false>, kBThreads),
                                   resident_ctas(bn_plane_bwdx_kernel<true, true>, kBThreads)};
        const int split = pick_split(C, N, reduce ? kMaxBnSplit : (1 << 20),
                                     kNumSM * occ[(relu ? 2 : 0) + (reduce ? 1 : 0)]);
        const unsigned grid = (unsigned)C * split;
        if (relu && reduce) bn_plane_bwdx_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, N, C, HW, split, P, dwdb, ws);
        else if (relu) bn_plane_bwdx_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, N, C, HW, split, P, dwdb, ws);
        else if (reduce) bn_plane_bwdx_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, N, C, HW, split, P, dwdb, ws);
        else bn_plane_bwdx_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, N, C, HW, split, P, dwdb, ws);
    } else {
        BnGeom G; int vec; char why[128];
        static const int per_sm = resident_ctas(bn_group_bwdx_kernel<4, true, true>, kBThreads);
        if (make_geom(N, C, HW, vec_ok, kNumSM * per_sm, G, vec, why, sizeof(why)) != OODFQ_OK)
            return fail(OODFQ_EINVAL, "bn_eval_backward: %s", why);
        const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
        if (vec == 4) {
            if (relu && reduce) bn_group_bwdx_kernel<4, true, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else if (relu) bn_group_bwdx_kernel<4, true, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else if (reduce) bn_group_bwdx_kernel<4, false, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else bn_group_bwdx_kernel<4, false, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
        } else {
            if (relu && reduce) bn_group_bwdx_kernel<1, true, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else if (relu) bn_group_bwdx_kernel<1, true, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else if (reduce) bn_group_bwdx_kernel<1, false, true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
            else bn_group_bwdx_kernel<1, false, false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, dwdb, ws);
        }
    }
    count_launch();
    return check_launch("bn_eval_backward");
}

extern "C" int oodfq_bn_eval_tap_backward(const float* x, const float* grad_y, float* grad_x, int N, int C, long long HW,
                                          const float* weight, const float* bias, const float* running_mean,
                                          const float* running_var, float eps, int flags, const float* mean,
                                          const float* gmean, const float* gvar, double count, const float* gscale,
                                          oodfq_stream_t stream) {
    if (!x || !grad_y || !grad_x || !running_mean || !running_var || !mean || !gmean || !gvar)
        return fail(OODFQ_EINVAL, "bn_eval_tap_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0 || !(count > 0)) return fail(OODFQ_EINVAL, "bn_eval_tap_backward: empty tensor");
    if (!(flags & OODFQ_BN_NHWC) || (C % 4) != 0 || !aligned16(x) || !aligned16(grad_y) || !aligned16(grad_x))
        return fail(OODFQ_EINVAL, "bn_eval_tap_backward: channels_last tensors with C %% 4 == 0 and 16-byte alignment only");
    const BnParams P{weight, bias, running_mean, running_var, eps};
    const NhwcGeom G = make_nhwc((long long)N * HW, C);
    const bool relu = flags & OODFQ_BN_RELU;
    static const int occ[2] = {resident_ctas(bn_nhwc_bwdx_tap_kernel<false>, kBThreads),
                               resident_ctas(bn_nhwc_bwdx_tap_kernel<true>, kBThreads)};
    long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
    const long long cap = (long long)kNumSM * occ[relu ? 1 : 0];
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
    const float ic = (float)(1.0 / count);
    cudaStream_t st = (cudaStream_t)stream;
    if (relu) bn_nhwc_bwdx_tap_kernel<true><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, mean, gmean, gvar, ic, gscale);
    else bn_nhwc_bwdx_tap_kernel<false><<<grid, kBThreads, 0, st>>>(x, grad_y, grad_x, G, P, mean, gmean, gvar, ic, gscale);
    count_launch();
    return check_launch("bn_eval_tap_backward");
}
