// Per-channel statistics of a BatchNorm input, the BN-statistics loss and its backward.
//
// Replaces the forward hook of the reference (trainer_direct.py:388-397,
// data_generate/distill_data.py:69-78: input.mean([0,2,3]) and
// input.var([0,2,3], unbiased=False), each a multi-pass ATen reduction with an
// autograd tape behind it) and the loss around it (trainer_direct.py:473-486,
// distill_data.py:252-265).
//
//   forward  : one read of x (4 B/elem) -> S1_c = sum(x - shift_c), S2_c = sum((x - shift_c)^2)
//              as fp64.  shift = the BN running mean: identical on every rank, so the sums
//              are additive across GPUs (one packed all-reduce).  Numerically each CTA
//              accumulates in fp32 around a LOCAL pivot (a sample of the channel it is
//              reading, so |x - pivot| ~ sigma) and re-bases its partial onto the shift in
//              fp64; nothing ever subtracts two large fp32 numbers.  Optionally the same
//              read also emits the fake-quantised tensor (north_star (b)).
//   loss     : one tiny kernel over the packed sums of all layers (after the optional NCCL
//              all-reduce) -> loss, mean, var, dL/dmean, dL/dvar.
//   backward : grad_x = grad_in + g*(gmean_c/M + gvar_c*2(x-mean_c)/M): one read of x
//              (+ grad_in), one write: 8 or 12 B/elem instead of an autograd tape.
//
// Two decompositions of the NCHW tensor:
//   plane kernels  (H*W >= 1024, 16-byte aligned): a CTA owns one channel and a subset of the
//              batch index and simply streams its planes -- 8 accumulators per thread.
//   group kernels  (small planes: 7x7, 14x14 ...; or unaligned): the tensor is [N][C*HW]; a CTA
//              owns a contiguous SPAN of the inner axis covering several whole channels (or a
//              chunk of one) and a subset of n.  A thread keeps the same offsets inside the span
//              for every n, so the channel of each of its elements is fixed and the sums stay in
//              registers; fully coalesced even for 49-element planes.
// Reductions are ordered (deterministic): fixed shuffle trees, per-CTA partials folded in index
// order by the last CTA of each channel (group); no floating-point atomics.
//
// Roofline: HBM for all kernels here.
#include "bn_geom.cuh"

namespace oodfq {

// re-base (count, S1, S2) taken around `pivot` onto `shift`, in fp64
__device__ __forceinline__ void rebase(double cnt, float t1, float t2, float pivot, float shift,
                                       double& s1, double& s2) {
    const double d = (double)pivot - (double)shift;
    s1 = (double)t1 + cnt * d;
    s2 = (double)t2 + 2.0 * d * (double)t1 + cnt * d * d;
}

// =============================================================================== plane kernels
template <bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_plane_stats_kernel(const float* __restrict__ x, int N, int C, long long HW, int split,
                      const float* __restrict__ shift, double* __restrict__ sums, float* __restrict__ y,
                      const float* __restrict__ fq_lo, const float* __restrict__ fq_hi, int fq_k, Workspace* ws) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    __shared__ float r1[kBThreads / 32], r2[kBThreads / 32];
    __shared__ int s_last;
    const int c = blockIdx.x / split, sp = blockIdx.x % split;
    QParams qp;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    const float pivot = __ldg(x + ((long long)sp * C + c) * HW);
    const int n4 = (int)(HW >> 2);
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    int iters = 0;
    for (int n = sp; n < N; n += split, ++iters) {
        const long long base = ((long long)n * C + c) * HW;
        const float4* p = reinterpret_cast<const float4*>(x + base);
        float4* q = QUANT ? reinterpret_cast<float4*>(y + base) : nullptr;
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
                    float d;
                    d = v[u].x - pivot; a1[0] += d; a2[0] = fmaf(d, d, a2[0]);
                    d = v[u].y - pivot; a1[1] += d; a2[1] = fmaf(d, d, a2[1]);
                    d = v[u].z - pivot; a1[2] += d; a2[2] = fmaf(d, d, a2[2]);
                    d = v[u].w - pivot; a1[3] += d; a2[3] = fmaf(d, d, a2[3]);
                    if (QUANT)
                        st_out(q + i, make_float4(fake_quant_lut(v[u].x, qp, lut, qh, qmask),
                                                  fake_quant_lut(v[u].y, qp, lut, qh, qmask),
                                                  fake_quant_lut(v[u].z, qp, lut, qh, qmask),
                                                  fake_quant_lut(v[u].w, qp, lut, qh, qmask)));
                }
            }
        }
    }
    float t1 = warp_sum((a1[0] + a1[1]) + (a1[2] + a1[3]));
    float t2 = warp_sum((a2[0] + a2[1]) + (a2[2] + a2[3]));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { r1[warp] = t1; r2[warp] = t2; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kBThreads / 32; ++w) { t1 += r1[w]; t2 += r2[w]; }
        double s1, s2;
        rebase((double)iters * (double)HW, t1, t2, pivot, shift ? __ldg(shift + c) : 0.f, s1, s2);
        double* pp = ws->bn_partial + ((size_t)sp * C + c) * 2;
        pp[0] = s1;
        pp[1] = s2;
        __threadfence();
        s_last = (atomicAdd(&ws->bn_ticket[c], 1) == split - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp == 0) fold_partials(ws->bn_partial, C, c, split, lane, sums);
    if (threadIdx.x == 0) ws->bn_ticket[c] = 0;
}

__global__ void __launch_bounds__(kBThreads)
bn_plane_bwd_kernel(const float* __restrict__ x, const float* grad_in, float* grad_x, int N, int C,
                    long long HW, int split, const float* __restrict__ mean, const float* __restrict__ gmean,
                    const float* __restrict__ gvar, float inv_count, const float* __restrict__ gscale) {
    const int c = blockIdx.x / split, sp = blockIdx.x % split;
    const float gs = gscale ? __ldg(gscale) : 1.0f;
    const float ca = gs * 2.0f * __ldg(gvar + c) * inv_count;
    const float cb = gs * __ldg(gmean + c) * inv_count;
    const float mu = __ldg(mean + c);
    const int n4 = (int)(HW >> 2);
    for (int n = sp; n < N; n += split) {
        const long long base = ((long long)n * C + c) * HW;
        const float4* p = reinterpret_cast<const float4*>(x + base);
        const float4* gi = grad_in ? reinterpret_cast<const float4*>(grad_in + base) : nullptr;
        float4* go = reinterpret_cast<float4*>(grad_x + base);
        for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * kBThreads) {
            float4 v[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) {
                    v[u] = ld_stream(p + i);
                    if (gi) g[u] = gi[i];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * kBThreads;
                if (i < n4) {
                    float4 r = make_float4(fmaf(ca, v[u].x - mu, cb), fmaf(ca, v[u].y - mu, cb),
                                           fmaf(ca, v[u].z - mu, cb), fmaf(ca, v[u].w - mu, cb));
                    if (gi) { r.x += g[u].x; r.y += g[u].y; r.z += g[u].z; r.w += g[u].w; }
                    st_out(go + i, r);
                }
            }
        }
    }
}

// =============================================================================== group kernels
// One vector slot per thread (fixed offset inside the span, hence fixed channels), kDepth rows of
// the batch in flight at once: the same bytes in flight as a wide tile, at a quarter of the
// per-thread state (pivots / coefficients), so 5-6 CTAs stay resident per SM.
template <int VEC, bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_group_stats_kernel(const float* __restrict__ x, const BnGeom G, const float* __restrict__ shift,
                      double* __restrict__ sums, float* __restrict__ y, const float* __restrict__ fq_lo,
                      const float* __restrict__ fq_hi, int fq_k, Workspace* ws) {
    __shared__ float s1[kBThreads * VEC];
    __shared__ float s2[kBThreads * VEC];
    __shared__ float lut[QUANT ? kLutMax : 1];
    __shared__ int s_last;

    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;
    const int hw = (int)((G.cg > 1) ? G.HW : 0x7fffffff);   // group mode: planes are small

    QParams qp;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }

    // this thread owns elements e0 .. e0+VEC-1 of the span, for every n of the CTA's subset.
    // pivot of an element = first element of its channel (or chunk) in the CTA's first row.
    const int e0 = threadIdx.x * VEC;
    const bool active = e0 < len;
    const float* x0 = x + sp * row + off;
    float a1[VEC], a2[VEC], pv[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        a1[j] = 0.f; a2[j] = 0.f;
        pv[j] = active ? __ldg(x0 + ((e0 + j) / hw) * hw) : 0.f;
    }

    int iters = 0;
    for (int n = sp; n < G.N; n += kDepth * G.split) {
        float v[kDepth][VEC];
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (active && nn < G.N) load_vec<VEC>(x + nn * row + off + e0, v[d]);
        }
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (nn < G.N) {
                ++iters;
                if (active) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        float dlt = v[d][j] - pv[j];
                        a1[j] += dlt;
                        a2[j] = fmaf(dlt, dlt, a2[j]);
                    }
                    if (QUANT) {
                        float r[VEC];
#pragma unroll
                        for (int j = 0; j < VEC; ++j) r[j] = fake_quant_lut(v[d][j], qp, lut, qh, qmask);
                        store_vec<VEC>(y + nn * row + off + e0, r);
                    }
                }
            }
        }
    }

    // ordered reduction: registers -> shared memory -> one warp per channel -> fp64 partial
    if (active) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s1[e0 + j] = a1[j]; s2[e0 + j] = a2[j]; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = (G.cg > 1) ? min(G.cg, G.C - c0) : 1;
    const int part = ck * G.split + sp;               // partial slot of this CTA
    const int nparts = per_group;
    for (int c = warp; c < nch; c += kBThreads / 32) {
        int eb = (G.cg > 1) ? (int)(c * G.HW) : 0;
        int ee = (G.cg > 1) ? (int)((c + 1) * G.HW) : len;
        float t1 = 0.f, t2 = 0.f;
        for (int e = eb + lane; e < ee; e += 32) { t1 += s1[e]; t2 += s2[e]; }
        t1 = warp_sum(t1);
        t2 = warp_sum(t2);
        if (lane == 0) {
            double d1, d2;
            rebase((double)iters * (double)(ee - eb), t1, t2, __ldg(x0 + eb), shift ? __ldg(shift + c0 + c) : 0.f,
                   d1, d2);
            double* p = ws->bn_partial + ((size_t)part * G.C + (c0 + c)) * 2;
            p[0] = d1;
            p[1] = d2;
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
    for (int c = warp; c < nch; c += kBThreads / 32) fold_partials(ws->bn_partial, G.C, c0 + c, nparts, lane, sums);
    if (threadIdx.x == 0) ws->bn_ticket[g] = 0;
}

template <int VEC>
__global__ void __launch_bounds__(kBThreads)
bn_group_bwd_kernel(const float* __restrict__ x, const float* grad_in, float* grad_x, const BnGeom G,
                    const float* __restrict__ mean, const float* __restrict__ gmean,
                    const float* __restrict__ gvar, float inv_count, const float* __restrict__ gscale) {
    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;
    const int hw = (int)((G.cg > 1) ? G.HW : 0x7fffffff);
    const float gs = gscale ? __ldg(gscale) : 1.0f;
    const int e0 = threadIdx.x * VEC;
    if (e0 >= len) return;

    float ca[VEC], cb[VEC], mu[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        int c = c0 + (e0 + j) / hw;
        ca[j] = gs * 2.0f * __ldg(gvar + c) * inv_count;
        cb[j] = gs * __ldg(gmean + c) * inv_count;
        mu[j] = __ldg(mean + c);
    }
    for (int n = sp; n < G.N; n += kDepth * G.split) {
        float v[kDepth][VEC], gi[kDepth][VEC];
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (nn < G.N) {
                const long long at = nn * row + off + e0;
                load_vec<VEC>(x + at, v[d]);
                if (grad_in) {
                    if (VEC == 4) {
                        float4 t = *reinterpret_cast<const float4*>(grad_in + at);
                        gi[d][0] = t.x; gi[d][1 % VEC] = t.y; gi[d][2 % VEC] = t.z; gi[d][3 % VEC] = t.w;
                    } else {
                        gi[d][0] = grad_in[at];
                    }
                }
            }
        }
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int nn = n + d * G.split;
            if (nn < G.N) {
                float r[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float t = fmaf(ca[j], v[d][j] - mu[j], cb[j]);
                    r[j] = grad_in ? gi[d][j] + t : t;
                }
                store_vec<VEC>(grad_x + nn * row + off + e0, r);
            }
        }
    }
}

// =============================================================================== NHWC kernels
// HEAD: what the optional output y is -- 0: fakequant(x) (QuantAct with channel statistics, north_star (b)),
// 1: [fakequant](BN_eval(x)), 2: [fakequant](relu(BN_eval(x))): the fused eval-mode BatchNorm (bn_fused.cu) whose INPUT is
// tapped by the BN-statistics loss, statistics and BatchNorm from ONE read of x (8 instead of 4 + 8 B/elem).
template <bool QUANT, int HEAD = 0>
__global__ void __launch_bounds__(kBThreads)
bn_nhwc_stats_kernel(const float* __restrict__ x, const NhwcGeom G, const float* __restrict__ shift,
                     float* __restrict__ y, const float* __restrict__ fq_lo, const float* __restrict__ fq_hi,
                     int fq_k, double* __restrict__ part, const BnParams P = BnParams{nullptr, nullptr, nullptr, nullptr, 0.f}) {
    __shared__ double dred[2 * kBThreads];
    __shared__ float lut[QUANT ? kLutMax : 1];
    QParams qp;
    float lowc = 0.0f;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        lowc = relu_lower_bound(qp);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
    }
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    const bool active = (int)threadIdx.x < G.lanes_r * wcols;
    const int rsub = threadIdx.x / wcols;
    const long long rstep = (long long)G.lanes_r * gridDim.x;
    const long long r0 = (long long)blockIdx.x * G.lanes_r + rsub;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        __syncthreads();                       // LUT ready / `dred` free again
        const int col = cb * kBThreads + threadIdx.x % wcols;
        const bool on = active && col < G.cols && r0 < G.R;
        float pv[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
        float ha[4] = {1.f, 1.f, 1.f, 1.f}, hb[4] = {0.f, 0.f, 0.f, 0.f};
        long long cnt = 0;
        if (HEAD && on) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { float invstd; affine_of(P, 4 * col + j, ha[j], hb[j], invstd); }
        }
        if (on) {
            // pivot = this thread's first row: a sample of each of its 4 channels
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(x) + r0 * G.cols + col);
            pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
            for (long long r = r0; r < G.R; r += kDepth * rstep) {
                float4 v[kDepth];
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const long long rr = r + d * rstep;
                    if (rr < G.R) v[d] = ld_stream(reinterpret_cast<const float4*>(x) + rr * G.cols + col);
                }
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const long long rr = r + d * rstep;
                    if (rr < G.R) {
                        ++cnt;
                        const float xs[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float dl = xs[j] - pv[j];
                            a1[j] += dl;
                            a2[j] = fmaf(dl, dl, a2[j]);
                        }
                        if (HEAD) {
                            float z, o[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[j] = head<HEAD == 2, QUANT>(xs[j], ha[j], hb[j], qp, lowc, lut, qh, qmask, z);
                            st_out(reinterpret_cast<float4*>(y) + rr * G.cols + col, make_float4(o[0], o[1], o[2], o[3]));
                        } else if (QUANT) {
                            st_out(reinterpret_cast<float4*>(y) + rr * G.cols + col,
                                   make_float4(fake_quant_lut(xs[0], qp, lut, qh, qmask), fake_quant_lut(xs[1], qp, lut, qh, qmask),
                                               fake_quant_lut(xs[2], qp, lut, qh, qmask), fake_quant_lut(xs[3], qp, lut, qh, qmask)));
                        }
                    }
                }
            }
        }
        // every thread re-bases its own (count, S1, S2) onto the shift in fp64; the CTA's row-lanes are then
        // folded in lane order through shared memory, one channel of the column at a time
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double d1 = 0.0, d2 = 0.0;
            if (on) rebase((double)cnt, a1[j], a2[j], pv[j], shift ? __ldg(shift + 4 * col + j) : 0.f, d1, d2);
            __syncthreads();
            dred[threadIdx.x] = d1;
            dred[kBThreads + threadIdx.x] = d2;
            __syncthreads();
            if (active && col < G.cols && rsub == 0) {
                const int lc = threadIdx.x % wcols;
                double t1 = 0.0, t2 = 0.0;
                for (int l = 0; l < G.lanes_r; ++l) { t1 += dred[l * wcols + lc]; t2 += dred[kBThreads + l * wcols + lc]; }
                double* p = part + ((size_t)blockIdx.x * G.C + 4 * col + j) * 2;
                p[0] = t1;
                p[1] = t2;
            }
        }
    }
}

__global__ void __launch_bounds__(kBThreads)
bn_nhwc_bwd_kernel(const float* __restrict__ x, const float* grad_in, float* grad_x, const NhwcGeom G,
                   const float* __restrict__ mean, const float* __restrict__ gmean, const float* __restrict__ gvar,
                   float inv_count, const float* __restrict__ gscale) {
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    if ((int)threadIdx.x >= G.lanes_r * wcols) return;
    const int rsub = threadIdx.x / wcols;
    const float gs = gscale ? __ldg(gscale) : 1.0f;
    const long long rstep = (long long)G.lanes_r * gridDim.x;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + threadIdx.x % wcols;
        if (col >= G.cols) continue;
        float ca[4], cbv[4], mu[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
                    if (grad_in) g[d] = reinterpret_cast<const float4*>(grad_in)[rr * G.cols + col];
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const long long rr = r + d * rstep;
                if (rr < G.R) {
                    float4 o = make_float4(fmaf(ca[0], v[d].x - mu[0], cbv[0]), fmaf(ca[1], v[d].y - mu[1], cbv[1]),
                                           fmaf(ca[2], v[d].z - mu[2], cbv[2]), fmaf(ca[3], v[d].w - mu[3], cbv[3]));
                    if (grad_in) { o.x += g[d].x; o.y += g[d].y; o.z += g[d].z; o.w += g[d].w; }
                    st_out(reinterpret_cast<float4*>(grad_x) + rr * G.cols + col, o);
                }
            }
        }
    }
}

// =============================================================================== small kernels
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ shift, int C,
                                   double inv_count, float* mean, float* var) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m1 = sums[c] * inv_count;
    double m2 = sums[C + c] * inv_count;
    mean[c] = (float)((shift ? (double)shift[c] : 0.0) + m1);
    var[c] = (float)(m2 - m1 * m1);
}

constexpr int kMaxLayers = 160;
struct BnsLayers {
    int off[kMaxLayers + 1];
    double inv_count[kMaxLayers];
    int L;
};

// single CTA: the packed arrays hold a few thousand channels at most
__global__ void __launch_bounds__(1024)
bns_loss_kernel(const double* __restrict__ sums, const float* __restrict__ shift,
                const float* __restrict__ run_mean, const float* __restrict__ run_var,
                const __grid_constant__ BnsLayers lay, float* loss3, float* mean, float* var,
                float* gmean, float* gvar) {
    __shared__ double r1[32], r2[32];
    const int Ctot = lay.off[lay.L];
    double lm = 0.0, lv = 0.0;
    for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
        int lo = 0, hi = lay.L - 1;           // layer of channel c
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (lay.off[mid] <= c) lo = mid; else hi = mid - 1;
        }
        const double ic = lay.inv_count[lo];
        // layer l keeps its sums together: [2*off_l, 2*off_l + C_l) = S1, then C_l values of S2
        const int Cl = lay.off[lo + 1] - lay.off[lo];
        const double* ls = sums + 2 * lay.off[lo];
        double m1 = ls[c - lay.off[lo]] * ic;
        double m2 = ls[Cl + c - lay.off[lo]] * ic;
        const double invC = 1.0 / (double)Cl;
        float mu = (float)((shift ? (double)shift[c] : 0.0) + m1);
        float vr = (float)(m2 - m1 * m1);
        double dm = (double)mu - (double)run_mean[c];
        double dv = (double)vr - (double)run_var[c];
        lm += dm * dm * invC;
        lv += dv * dv * invC;
        mean[c] = mu;
        var[c] = vr;
        gmean[c] = (float)(2.0 * dm * invC / (double)lay.L);
        gvar[c] = (float)(2.0 * dv * invC / (double)lay.L);
    }
    lm = warp_sum(lm);
    lv = warp_sum(lv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { r1[warp] = lm; r2[warp] = lv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r1[w]; b += r2[w]; }
        loss3[0] = (float)((a + b) / (double)lay.L);
        loss3[1] = (float)(a / (double)lay.L);
        loss3[2] = (float)(b / (double)lay.L);
    }
}

// =============================================================================== host side
}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_bn_stats_forward(const float* x, int N, int C, long long HW, const float* shift,
                                      double* sums, float* y, const float* fq_lo, const float* fq_hi,
                                      int fq_k, int flags, void* workspace, oodfq_stream_t stream) {
    if (!x || !sums || !workspace) return fail(OODFQ_EINVAL, "bn_stats_forward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "bn_stats_forward: empty tensor (N=%d C=%d HW=%lld)", N, C, HW);
    if (C > kMaxBnChannels) return fail(OODFQ_EINVAL, "bn_stats_forward: C=%d exceeds %d", C, kMaxBnChannels);
    if (y && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8)) return fail(OODFQ_EINVAL, "bn_stats_forward: fused fake-quant needs a range and k in [1,8]");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const bool vec_ok = aligned16(x) && (!y || aligned16(y));
    if (flags & OODFQ_BN_NHWC) {
        if (!vec_ok || (C % 4) != 0) return fail(OODFQ_EINVAL, "bn_stats_forward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        const NhwcGeom G = make_nhwc((long long)N * HW, C);
        static const int occ[2] = {resident_ctas(bn_nhwc_stats_kernel<false>, kBThreads),
                                   resident_ctas(bn_nhwc_stats_kernel<true>, kBThreads)};
        long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
        long long cap = (long long)kNumSM * occ[y ? 1 : 0];
        const long long table = (long long)kMaxBnSplit * kMaxBnChannels / C;
        if (cap > table) cap = table;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        double* part = fold_target(ws->bn_partial, C, (int)grid);          // workspace, or its own region (fold.cu)
        if (y) bn_nhwc_stats_kernel<true><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part);
        else bn_nhwc_stats_kernel<false><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part);
        count_launch();
        int rc = check_launch("bn_stats_forward");
        if (rc != OODFQ_OK) return rc;
        return fold_finish(part, ws->bn_partial, C, (int)grid, sums, st);
    }
    if (plane_ok(HW, vec_ok)) {
        static const int per_sm_q = resident_ctas(bn_plane_stats_kernel<true>, kBThreads);
        static const int per_sm = resident_ctas(bn_plane_stats_kernel<false>, kBThreads);
        const int split = pick_split(C, N, kMaxBnSplit, kNumSM * (y ? per_sm_q : per_sm));
        const unsigned grid = (unsigned)C * split;
        if (y) bn_plane_stats_kernel<true><<<grid, kBThreads, 0, st>>>(x, N, C, HW, split, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        else bn_plane_stats_kernel<false><<<grid, kBThreads, 0, st>>>(x, N, C, HW, split, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        count_launch();
        return check_launch("bn_stats_forward");
    }
    BnGeom G; int vec; char why[128];
    static const int per_sm_g = resident_ctas(bn_group_stats_kernel<4, true>, kBThreads);
    if (make_geom(N, C, HW, vec_ok, kNumSM * per_sm_g, G, vec, why, sizeof(why)) != OODFQ_OK) return fail(OODFQ_EINVAL, "bn_stats_forward: %s", why);
    const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
    if (vec == 4) {
        if (y) bn_group_stats_kernel<4, true><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        else bn_group_stats_kernel<4, false><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
    } else {
        if (y) bn_group_stats_kernel<1, true><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        else bn_group_stats_kernel<1, false><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
    }
    count_launch();
    return check_launch("bn_stats_forward");
}

extern "C" int oodfq_bn_stats_finalize(const double* sums, const float* shift, int C, double count,
                                       float* mean, float* var, oodfq_stream_t stream) {
    if (!sums || !mean || !var) return fail(OODFQ_EINVAL, "bn_stats_finalize: null pointer");
    if (C <= 0 || !(count > 0)) return fail(OODFQ_EINVAL, "bn_stats_finalize: C=%d count=%g", C, count);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, shift, C, 1.0 / count, mean, var);
    count_launch();
    return check_launch("bn_stats_finalize");
}

extern "C" int oodfq_bns_loss(const double* sums, const float* shift, const float* run_mean,
                              const float* run_var, const int* ch_off_host, const double* counts_host,
                              int L, float* loss3, float* mean, float* var, float* gmean, float* gvar,
                              oodfq_stream_t stream) {
    if (!sums || !run_mean || !run_var || !ch_off_host || !counts_host || !loss3 || !mean || !var || !gmean || !gvar)
        return fail(OODFQ_EINVAL, "bns_loss: null pointer");
    if (L < 1 || L > kMaxLayers) return fail(OODFQ_EINVAL, "bns_loss: L=%d outside [1,%d]", L, kMaxLayers);
    BnsLayers lay;
    lay.L = L;
    for (int l = 0; l <= L; ++l) lay.off[l] = ch_off_host[l];
    for (int l = 0; l < L; ++l) {
        if (!(counts_host[l] > 0) || ch_off_host[l + 1] <= ch_off_host[l])
            return fail(OODFQ_EINVAL, "bns_loss: layer %d has no elements", l);
        lay.inv_count[l] = 1.0 / counts_host[l];
    }
    if (ch_off_host[0] != 0) return fail(OODFQ_EINVAL, "bns_loss: ch_off[0] must be 0");
    bns_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sums, shift, run_mean, run_var, lay, loss3, mean, var, gmean, gvar);
    count_launch();
    return check_launch("bns_loss");
}

extern "C" int oodfq_bn_stats_backward(const float* x, const float* grad_in, float* grad_x, int N, int C,
                                       long long HW, const float* mean, const float* gmean,
                                       const float* gvar, double count, const float* gscale, int flags,
                                       oodfq_stream_t stream) {
    if (!x || !grad_x || !mean || !gmean || !gvar) return fail(OODFQ_EINVAL, "bn_stats_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0 || !(count > 0)) return fail(OODFQ_EINVAL, "bn_stats_backward: empty tensor");
    cudaStream_t st = (cudaStream_t)stream;
    const float ic = (float)(1.0 / count);
    const bool vec_ok = aligned16(x) && aligned16(grad_x) && (!grad_in || aligned16(grad_in));
    if (flags & OODFQ_BN_NHWC) {
        if (!vec_ok || (C % 4) != 0) return fail(OODFQ_EINVAL, "bn_stats_backward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        const NhwcGeom G = make_nhwc((long long)N * HW, C);
        static const int per_sm = resident_ctas(bn_nhwc_bwd_kernel, kBThreads);
        long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
        long long cap = (long long)kNumSM * per_sm;
        const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        bn_nhwc_bwd_kernel<<<grid, kBThreads, 0, st>>>(x, grad_in, grad_x, G, mean, gmean, gvar, ic, gscale);
        count_launch();
        return check_launch("bn_stats_backward");
    }
    if (plane_ok(HW, vec_ok)) {
        static const int per_sm = resident_ctas(bn_plane_bwd_kernel, kBThreads);
        const int split = pick_split(C, N, 1 << 20, kNumSM * per_sm);
        bn_plane_bwd_kernel<<<(unsigned)C * split, kBThreads, 0, st>>>(x, grad_in, grad_x, N, C, HW, split, mean, gmean, gvar, ic, gscale);
        count_launch();
        return check_launch("bn_stats_backward");
    }
    BnGeom G; int vec; char why[128];
    static const int per_sm_g = resident_ctas(bn_group_bwd_kernel<4>, kBThreads);
    if (make_geom(N, C, HW, vec_ok, kNumSM * per_sm_g, G, vec, why, sizeof(why)) != OODFQ_OK) return fail(OODFQ_EINVAL, "bn_stats_backward: %s", why);
    const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
    if (vec == 4) bn_group_bwd_kernel<4><<<grid, kBThreads, 0, st>>>(x, grad_in, grad_x, G, mean, gmean, gvar, ic, gscale);
    else bn_group_bwd_kernel<1><<<grid, kBThreads, 0, st>>>(x, grad_in, grad_x, G, mean, gmean, gvar, ic, gscale);
    count_launch();
    return check_launch("bn_stats_backward");
}

extern "C" int oodfq_bn_eval_stats_forward(const float* x, float* y, int N, int C, long long HW, const float* weight,
                                           const float* bias, const float* running_mean, const float* running_var,
                                           float eps, int flags, const float* fq_lo, const float* fq_hi, int fq_k,
                                           const float* shift, double* sums, void* workspace, oodfq_stream_t stream) {
    if (!x || !y || !running_mean || !running_var || !sums || !workspace)
        return fail(OODFQ_EINVAL, "bn_eval_stats_forward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "bn_eval_stats_forward: empty tensor");
    if (C > kMaxBnChannels) return fail(OODFQ_EINVAL, "bn_eval_stats_forward: C=%d exceeds %d", C, kMaxBnChannels);
    if (!(flags & OODFQ_BN_NHWC) || (C % 4) != 0 || !aligned16(x) || !aligned16(y))
        return fail(OODFQ_EINVAL, "bn_eval_stats_forward: channels_last tensors with C %% 4 == 0 and 16-byte alignment only");
    const bool relu = flags & OODFQ_BN_RELU, quant = flags & OODFQ_BN_QUANT;
    if (quant && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8))
        return fail(OODFQ_EINVAL, "bn_eval_stats_forward: fused fake-quant needs a range and k in [1,8]");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const BnParams P{weight, bias, running_mean, running_var, eps};
    const NhwcGeom G = make_nhwc((long long)N * HW, C);
    static const int occ[4] = {resident_ctas(bn_nhwc_stats_kernel<false, 1>, kBThreads), resident_ctas(bn_nhwc_stats_kernel<false, 2>, kBThreads),
                               resident_ctas(bn_nhwc_stats_kernel<true, 1>, kBThreads), resident_ctas(bn_nhwc_stats_kernel<true, 2>, kBThreads)};
    long long want = (G.R + (long long)G.lanes_r * kDepth - 1) / ((long long)G.lanes_r * kDepth);
    long long cap = (long long)kNumSM * occ[(quant ? 2 : 0) + (relu ? 1 : 0)];
    const long long table = (long long)kMaxBnSplit * kMaxBnChannels / C;
    if (cap > table) cap = table;
    const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
    double* part = fold_target(ws->bn_partial, C, (int)grid);              // workspace, or its own region (fold.cu)
    if (quant && relu) bn_nhwc_stats_kernel<true, 2><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part, P);
    else if (quant) bn_nhwc_stats_kernel<true, 1><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part, P);
    else if (relu) bn_nhwc_stats_kernel<false, 2><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part, P);
    else bn_nhwc_stats_kernel<false, 1><<<grid, kBThreads, 0, st>>>(x, G, shift, y, fq_lo, fq_hi, fq_k, part, P);
    count_launch();
    int rc = check_launch("bn_eval_stats_forward");
    if (rc != OODFQ_OK) return rc;
    return fold_finish(part, ws->bn_partial, C, (int)grid, sums, st);
}
