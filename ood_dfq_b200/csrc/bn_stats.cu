// Per-channel statistics of a BatchNorm input, the BN-statistics loss and its backward.
//
// Replaces the forward hook of the reference (trainer_direct.py:388-397,
// data_generate/distill_data.py:69-78: input.mean([0,2,3]) and
// input.var([0,2,3], unbiased=False), each a multi-pass ATen reduction with an
// autograd tape behind it) and the loss around it (trainer_direct.py:473-486,
// distill_data.py:252-265).
//
//   forward  : one read of x (4 B/elem) -> shifted sums S1_c = sum(x - shift_c),
//              S2_c = sum((x - shift_c)^2).  shift = the BN running mean, which
//              removes the cancellation of the naive sum-of-squares formula and,
//              being identical on every rank, keeps the sums additive across GPUs.
//              Optionally the same read also emits the fake-quantised tensor.
//   loss     : one tiny kernel over the packed sums of all layers (after the
//              optional NCCL all-reduce) -> loss, mean, var, dL/dmean, dL/dvar.
//   backward : grad_x = grad_in + g*(gmean_c/M + gvar_c*2(x-mean_c)/M): one read of
//              x (+ grad_in), one write: 8 or 12 B/elem instead of an autograd tape.
//
// Decomposition (all three big kernels): the NCHW tensor is [N][C*HW]; a CTA owns a
// contiguous SPAN of that inner axis -- several whole channels when planes are
// small (7x7, 4x4), or a chunk of one plane when they are large -- and a subset of
// the batch index n.  A thread keeps the same offsets inside the span for every n,
// so the channel of each of its elements is fixed and sums stay in registers.
// The cross-thread reduction is ordered (deterministic): per-slot sums go to shared
// memory, each channel's slice is folded by one warp, per-CTA partials are folded
// in index order by the last CTA of each channel group.
//
// Roofline: HBM for all three kernels.
#include <cstdio>

#include "common.cuh"

namespace oodfq {

constexpr int kBThreads = 256;
constexpr int kSlots = 4;                           // vector slots per thread and per n
constexpr int kSpanMax = kBThreads * kSlots * 4;    // 4096 elements = 16 KB per n and CTA

struct BnGeom {
    int N, C;
    long long HW;
    int cg;          // channels per group (1 when a plane is split into chunks)
    int groups;      // ceil(C / cg)
    int chunks;      // chunks per plane (1 when cg > 1)
    long long chunk_len;   // elements per chunk (multiple of VEC)
    int split;       // CTAs along n
};

// span of CTA (g, ch): element offset inside the [C*HW] row and its length
__device__ __forceinline__ void cta_span(const BnGeom& G, int g, int ck, long long& off, int& len, int& c0) {
    c0 = g * G.cg;
    if (G.cg > 1) {
        int nch = min(G.cg, G.C - c0);
        off = (long long)c0 * G.HW;
        len = (int)(nch * G.HW);
    } else {
        long long s = (long long)ck * G.chunk_len;
        long long e = min(G.HW, s + G.chunk_len);
        off = (long long)c0 * G.HW + s;
        len = (int)(e - s);
    }
}

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC], bool keep) {
    if (VEC == 4) {
        float4 t;
        if (keep) t = __ldg(reinterpret_cast<const float4*>(p));
        else t = ld_stream(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = keep ? __ldg(p) : ld_stream(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
    if (VEC == 4) st_out(reinterpret_cast<float4*>(p), make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
    else *p = v[0];
}

// ------------------------------------------------------------------------------ forward
template <int VEC, bool QUANT>
__global__ void __launch_bounds__(kBThreads)
bn_stats_kernel(const float* __restrict__ x, const BnGeom G, const float* __restrict__ shift,
                float* __restrict__ sums, float* __restrict__ y, const float* __restrict__ fq_lo,
                const float* __restrict__ fq_hi, int fq_k, Workspace* ws) {
    __shared__ float s1[kSpanMax / 4 * VEC];
    __shared__ float s2[kSpanMax / 4 * VEC];
    __shared__ int s_last;

    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;

    __shared__ float lut[QUANT ? kLutMax : 1];
    QParams qp;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }

    // fixed per-thread slots: element e = (threadIdx.x + u*kBThreads)*VEC + j of the span
    float a1[kSlots][VEC], a2[kSlots][VEC], sh[kSlots][VEC];
#pragma unroll
    for (int u = 0; u < kSlots; ++u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            a1[u][j] = 0.f; a2[u][j] = 0.f;
            int e = (threadIdx.x + u * kBThreads) * VEC + j;
            int c = (G.cg > 1) ? c0 + e / (int)G.HW : c0;
            sh[u][j] = (shift && e < len) ? __ldg(shift + c) : 0.f;
        }
    }

    for (int n = sp; n < G.N; n += G.split) {
        const float* xr = x + n * row + off;
        float v[kSlots][VEC];
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            int e = (threadIdx.x + u * kBThreads) * VEC;
            if (e < len) load_vec<VEC>(xr + e, v[u], /*keep=*/false);
        }
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            int e = (threadIdx.x + u * kBThreads) * VEC;
            if (e < len) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float d = v[u][j] - sh[u][j];
                    a1[u][j] += d;
                    a2[u][j] = fmaf(d, d, a2[u][j]);
                }
                if (QUANT) {
                    float r[VEC];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) r[j] = fake_quant_lut(v[u][j], qp, lut, qh, qmask);
                    store_vec<VEC>(y + n * row + off + e, r);
                }
            }
        }
    }

    // ordered reduction: slots -> shared memory -> one warp per channel
#pragma unroll
    for (int u = 0; u < kSlots; ++u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int e = (threadIdx.x + u * kBThreads) * VEC + j;
            if (e < len) { s1[e] = a1[u][j]; s2[e] = a2[u][j]; }
        }
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
            float* p = ws->bn_partial + ((size_t)part * G.C + (c0 + c)) * 2;
            p[0] = t1;
            p[1] = t2;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int t = atomicAdd(&ws->bn_ticket[g], 1);
        s_last = (t == nparts - 1);
    }
    __syncthreads();
    if (!s_last) return;

    // last CTA of this channel group: fold the partials in index order
    // (one warp per channel, lanes stride over the partials, fixed shuffle tree: deterministic)
    __threadfence();
    for (int c = warp; c < nch; c += kBThreads / 32) {
        double t1 = 0.0, t2 = 0.0;
        for (int p = lane; p < nparts; p += 32) {
            const float2 q = __ldcg(reinterpret_cast<const float2*>(ws->bn_partial + ((size_t)p * G.C + (c0 + c)) * 2));
            t1 += (double)q.x;
            t2 += (double)q.y;
        }
        t1 = warp_sum(t1);
        t2 = warp_sum(t2);
        if (lane == 0) {
            sums[c0 + c] = (float)t1;
            sums[G.C + c0 + c] = (float)t2;
        }
    }
    if (threadIdx.x == 0) ws->bn_ticket[g] = 0;
}

// ------------------------------------------------------------------------------ backward
template <int VEC>
__global__ void __launch_bounds__(kBThreads)
bn_stats_bwd_kernel(const float* __restrict__ x, const float* grad_in, float* grad_x, const BnGeom G,
                    const float* __restrict__ mean, const float* __restrict__ gmean,
                    const float* __restrict__ gvar, float inv_count, const float* __restrict__ gscale) {
    const int per_group = G.chunks * G.split;
    const int g = blockIdx.x / per_group;
    const int ck = (blockIdx.x % per_group) / G.split;
    const int sp = blockIdx.x % G.split;
    long long off; int len, c0;
    cta_span(G, g, ck, off, len, c0);
    const long long row = (long long)G.C * G.HW;
    const float gs = gscale ? __ldg(gscale) : 1.0f;

    float ca[kSlots][VEC], cb[kSlots][VEC], mu[kSlots][VEC];
#pragma unroll
    for (int u = 0; u < kSlots; ++u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            int e = (threadIdx.x + u * kBThreads) * VEC + j;
            int c = (G.cg > 1) ? c0 + e / (int)G.HW : c0;
            bool ok = e < len;
            ca[u][j] = ok ? gs * 2.0f * __ldg(gvar + c) * inv_count : 0.f;
            cb[u][j] = ok ? gs * __ldg(gmean + c) * inv_count : 0.f;
            mu[u][j] = ok ? __ldg(mean + c) : 0.f;
        }
    }
    for (int n = sp; n < G.N; n += G.split) {
        const long long base = n * row + off;
        float v[kSlots][VEC], gi[kSlots][VEC];
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            int e = (threadIdx.x + u * kBThreads) * VEC;
            if (e < len) {
                load_vec<VEC>(x + base + e, v[u], false);
                if (grad_in) {
                    if (VEC == 4) {
                        float4 t = *reinterpret_cast<const float4*>(grad_in + base + e);
                        gi[u][0] = t.x; gi[u][1 % VEC] = t.y; gi[u][2 % VEC] = t.z; gi[u][3 % VEC] = t.w;
                    } else {
                        gi[u][0] = grad_in[base + e];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kSlots; ++u) {
            int e = (threadIdx.x + u * kBThreads) * VEC;
            if (e < len) {
                float r[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float t = fmaf(ca[u][j], v[u][j] - mu[u][j], cb[u][j]);
                    r[j] = grad_in ? gi[u][j] + t : t;
                }
                store_vec<VEC>(grad_x + base + e, r);
            }
        }
    }
}

// ------------------------------------------------------------------------------ small kernels
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ shift, int C,
                                   double inv_count, float* mean, float* var) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m1 = (double)sums[c] * inv_count;
    double m2 = (double)sums[C + c] * inv_count;
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
bns_loss_kernel(const float* __restrict__ sums, const float* __restrict__ shift,
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
        const float* ls = sums + 2 * lay.off[lo];
        double m1 = (double)ls[c - lay.off[lo]] * ic;
        double m2 = (double)ls[Cl + c - lay.off[lo]] * ic;
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

// ------------------------------------------------------------------------------ host side
static int make_geom(int N, int C, long long HW, bool vec_ok, BnGeom& G, int& vec, char* why, size_t whyn) {
    G.N = N; G.C = C; G.HW = HW;
    // 128-bit access needs every span start and length to be a multiple of 4 elements
    vec = (vec_ok && ((long long)C * HW) % 4 == 0) ? 4 : 1;
    for (;;) {
        const long long span_max = (vec == 4) ? kSpanMax : kSpanMax / 4;
        if (HW <= span_max / 2) {            // several whole channels per CTA
            long long cg = span_max / HW;
            if (cg >= C) {
                cg = C;                      // one group starting at channel 0
            } else if (vec == 4 && HW % 4 != 0) {
                cg = (cg / 4) * 4;           // group starts stay 16-byte aligned
                if (cg == 0) { vec = 1; continue; }
            }
            G.cg = (int)cg;
            G.groups = (C + G.cg - 1) / G.cg;
            G.chunks = 1;
            G.chunk_len = (long long)G.cg * HW;
        } else {                             // one channel, plane cut into chunks
            if (vec == 4 && HW % 4 != 0) { vec = 1; continue; }
            long long chunks = (HW + span_max - 1) / span_max;
            long long cl = (HW + chunks - 1) / chunks;
            cl = ((cl + 3) / 4) * 4;
            if (cl > span_max) cl = span_max;
            chunks = (HW + cl - 1) / cl;
            G.cg = 1;
            G.groups = C;
            G.chunks = (int)chunks;
            G.chunk_len = cl;
        }
        break;
    }
    if (G.chunks > kMaxBnSplit) {
        snprintf(why, whyn, "plane of %lld elements needs %d chunks (max %d)", HW, G.chunks, kMaxBnSplit);
        return OODFQ_EINVAL;
    }
    // enough CTAs to fill the machine a few times, bounded by the partial table
    long long base = (long long)G.groups * G.chunks;
    long long want = ((long long)kNumSM * 8 + base - 1) / base;
    long long cap = kMaxBnSplit / G.chunks;
    if (want > cap) want = cap;
    if (want > N) want = N;
    if (want < 1) want = 1;
    G.split = (int)want;
    return OODFQ_OK;
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_bn_stats_forward(const float* x, int N, int C, long long HW, const float* shift,
                                      float* sums, float* y, const float* fq_lo, const float* fq_hi,
                                      int fq_k, void* workspace, oodfq_stream_t stream) {
    if (!x || !sums || !workspace) return fail(OODFQ_EINVAL, "bn_stats_forward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "bn_stats_forward: empty tensor (N=%d C=%d HW=%lld)", N, C, HW);
    if (C > kMaxBnChannels) return fail(OODFQ_EINVAL, "bn_stats_forward: C=%d exceeds %d", C, kMaxBnChannels);
    if (y && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8)) return fail(OODFQ_EINVAL, "bn_stats_forward: fused fake-quant needs a range and k in [1,8]");
    BnGeom G; int vec; char why[128];
    bool vec_ok = aligned16(x) && (!y || aligned16(y));
    if (make_geom(N, C, HW, vec_ok, G, vec, why, sizeof(why)) != OODFQ_OK) return fail(OODFQ_EINVAL, "bn_stats_forward: %s", why);
    const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    if (vec == 4) {
        if (y) bn_stats_kernel<4, true><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        else bn_stats_kernel<4, false><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
    } else {
        if (y) bn_stats_kernel<1, true><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
        else bn_stats_kernel<1, false><<<grid, kBThreads, 0, st>>>(x, G, shift, sums, y, fq_lo, fq_hi, fq_k, ws);
    }
    count_launch();
    return check_launch("bn_stats_forward");
}

extern "C" int oodfq_bn_stats_finalize(const float* sums, const float* shift, int C, double count,
                                       float* mean, float* var, oodfq_stream_t stream) {
    if (!sums || !mean || !var) return fail(OODFQ_EINVAL, "bn_stats_finalize: null pointer");
    if (C <= 0 || !(count > 0)) return fail(OODFQ_EINVAL, "bn_stats_finalize: C=%d count=%g", C, count);
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, shift, C, 1.0 / count, mean, var);
    count_launch();
    return check_launch("bn_stats_finalize");
}

extern "C" int oodfq_bns_loss(const float* sums, const float* shift, const float* run_mean,
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
                                       const float* gvar, double count, const float* gscale,
                                       oodfq_stream_t stream) {
    if (!x || !grad_x || !mean || !gmean || !gvar) return fail(OODFQ_EINVAL, "bn_stats_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0 || !(count > 0)) return fail(OODFQ_EINVAL, "bn_stats_backward: empty tensor");
    BnGeom G; int vec; char why[128];
    bool vec_ok = aligned16(x) && aligned16(grad_x) && (!grad_in || aligned16(grad_in));
    if (make_geom(N, C, HW, vec_ok, G, vec, why, sizeof(why)) != OODFQ_OK) return fail(OODFQ_EINVAL, "bn_stats_backward: %s", why);
    const unsigned grid = (unsigned)((long long)G.groups * G.chunks * G.split);
    cudaStream_t st = (cudaStream_t)stream;
    const float ic = (float)(1.0 / count);
    if (vec == 4) bn_stats_bwd_kernel<4><<<grid, kBThreads, 0, st>>>(x, grad_in, grad_x, G, mean, gmean, gvar, ic, gscale);
    else bn_stats_bwd_kernel<1><<<grid, kBThreads, 0, st>>>(x, grad_in, grad_x, G, mean, gmean, gvar, ic, gscale);
    count_launch();
    return check_launch("bn_stats_backward");
}
