// Space-to-depth re-layout of the image batch in front of the ImageNet stem convolution, channels_last.
//
// The first convolution of the ImageNet ResNets (3 -> 64 channels, 7x7, stride 2, padding 3; pytorchcv
// ResInitBlock behind ptcv_get_model, main_direct.py:380-397, wrapped by Quant_Conv2d :455-458) is the one
// layer cuDNN has no good kernel for: with 3 input channels it falls back to `implicit_gemm_indexed_wo_smem`
// at ~49 TFLOP/s and was 29 % of the fused step (profiles/r1_step_share_tail.txt: fprop 4.9 ms, wgrad 2.4 ms,
// dgrad 1.9 ms, layout helpers 3 ms).  A stride-2 KxK convolution equals a stride-1 ceil(K/2) x ceil(K/2)
// convolution over the 2x2 space-to-depth image:
//
//     xs[n, i, j, (s,t,c)] = x[n, 2i+s-p, 2j+t-p, c]   (zero outside the image),   i < (H+2p)/2, j < (W+2p)/2
//     w2[o, (s,t,c), a, b] = w[o, c, 2a+s, 2b+t]       (zero for 2a+s = K or 2b+t = K)
//     conv(x, w, stride 2, padding p)  ==  conv(xs, w2, stride 1, padding 0)          same products, same sums
//
// which cuDNN runs on its tensor-core implicit-GEMM path (12 channels, 4x4 taps): fprop 1.32 -> 0.72 ms,
// dgrad 1.99 -> 0.94 ms, wgrad 1.78 -> 1.37 ms at 256x3x224x224 (profiles/r1_exp_stem_conv.txt).  The
// convolution itself stays on cuDNN (BASELINE.json north_star); these two kernels are only the re-layout of
// its input (and of the input gradient on the way back), which in eager PyTorch is three passes (pad, permute,
// channels_last copy: 0.37 ms) and here one: 4 B read + 4 B written per element.  Roofline: HBM.
#include "common.cuh"

namespace oodfq {

struct S2dGeom {
    int N, H, W, C, pad, Hs, Ws;
    int CP;             // floats per xs pixel: 4*C, or more (zero channels behind the 4*C real ones, e.g. 16 for C = 3:
                        // cuDNN converts a 12-channel tensor before every use, a 16-channel one it takes as it is)
};

// generic channel counts: one thread per element (gather), runtime divisors
template <bool BWD>
__global__ void __launch_bounds__(256)
s2d_stem_kernel(const float* __restrict__ src, float* __restrict__ dst, const S2dGeom G, long long total) {
    const int c4 = G.CP, c2 = 2 * G.C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        if (!BWD) {
            // e indexes xs[n][i][j][ch]
            const int ch = (int)(e % c4);
            long long q = e / c4;
            if (ch >= 4 * G.C) { dst[e] = 0.0f; continue; }          // padding channel
            const int j = (int)(q % G.Ws); q /= G.Ws;
            const int i = (int)(q % G.Hs);
            const long long n = q / G.Hs;
            const int s = ch / c2, t = (ch / G.C) & 1, c = ch % G.C;
            const int h = 2 * i + s - G.pad, w = 2 * j + t - G.pad;
            float v = 0.0f;
            if (h >= 0 && h < G.H && w >= 0 && w < G.W) v = __ldg(src + ((n * G.H + h) * G.W + w) * G.C + c);
            dst[e] = v;
        } else {
            // e indexes gx[n][h][w][c]; every image pixel lives in exactly one xs element
            const int c = (int)(e % G.C);
            long long q = e / G.C;
            const int w = (int)(q % G.W); q /= G.W;
            const int h = (int)(q % G.H);
            const long long n = q / G.H;
            const int hp = h + G.pad, wp = w + G.pad;
            const int ch = ((hp & 1) * 2 + (wp & 1)) * G.C + c;
            dst[e] = __ldg(src + ((n * G.Hs + (hp >> 1)) * G.Ws + (wp >> 1)) * c4 + ch);
        }
    }
}

// Row form for the channel counts that matter (C = 3, and 1 / 4): for a fixed (image, xs row i, parity s) the
// s-half of every xs pixel is the input row h = 2i+s-pad, shifted by pad pixels and cut into runs of 2C floats:
//     xs_row[(q / 2C) * 4C + s*2C + q % 2C] = x_row[q - pad*C]        q in [0, 2C * Ws)
// A CTA step writes one whole destination row contiguously; all divisors are compile-time constants.
template <int C, bool BWD>
__global__ void __launch_bounds__(256)
s2d_stem_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const S2dGeom G) {
    const int shift = G.pad * C, wlen = G.W * C;
    if (!BWD) {
        // one xs row (n, i) per step: contiguous stores, loads from the two image rows 2i-pad and 2i+1-pad
        const long long rows = (long long)G.N * G.Hs;
        const int qn = 4 * C * G.Ws;
        for (long long ni = blockIdx.x; ni < rows; ni += gridDim.x) {
            const int i = (int)(ni % G.Hs);
            const long long n = ni / G.Hs;
            const int h0 = 2 * i - G.pad;
            const long long xrow0 = (n * G.H + h0) * (long long)wlen;
            float* out = dst + ni * (long long)qn;
            for (int q = threadIdx.x; q < qn; q += 256) {
                const int r = q % (4 * C), s = r / (2 * C);
                const int m = (q / (4 * C)) * (2 * C) + r % (2 * C) - shift;
                const int h = h0 + s;
                float v = 0.0f;
                if (h >= 0 && h < G.H && m >= 0 && m < wlen) v = __ldg(src + xrow0 + (long long)s * wlen + m);
                out[q] = v;
            }
        }
    } else {
        // one image row (n, h) per step: contiguous stores, loads in runs of 2C floats from xs row (h+pad)/2
        const long long rows = (long long)G.N * G.H;
        for (long long nh = blockIdx.x; nh < rows; nh += gridDim.x) {
            const int h = (int)(nh % G.H);
            const long long n = nh / G.H;
            const int hp = h + G.pad;
            const float* in = src + (n * G.Hs + (hp >> 1)) * (long long)(4 * C * G.Ws) + (hp & 1) * 2 * C;
            float* out = dst + nh * (long long)wlen;
            for (int m = threadIdx.x; m < wlen; m += 256) {
                const int q = m + shift;
                out[m] = __ldg(in + (q / (2 * C)) * (4 * C) + q % (2 * C));
            }
        }
    }
}

// TMA-staged row form (the default for C = 1 / 3 / 4 when rows are 16-byte multiples).  The scalar row kernel above
// keeps one 4-byte load per thread in flight and sat at 50 % of the HBM rate (profiles/r1 bench table).  Here a
// persistent CTA walks its xs rows through a ring of kS2dStages shared-memory stages: one thread issues the bulk
// copies of the NEXT rows' sources (cp.async.bulk global -> shared, completion on an mbarrier) while all threads
// permute the current source into the destination row inside shared memory (LDS / STS only), and the finished row
// leaves as one bulk store (shared -> global).  No data passes through registers on its way from or to HBM, and
// kS2dStages rows per CTA are in flight.
//   forward : source = the two image rows 2i-pad, 2i+1-pad (W*C floats each), destination = xs row (n, i)
//   backward: source = xs row (n, i), destination = the same two image rows of the input gradient
// Rows outside the image are neither copied nor stored; their elements read as zero.
constexpr int kS2dStages = 4;
constexpr int kS2dThreads = 128;

template <int C, bool BWD>
__global__ void __launch_bounds__(kS2dThreads)
s2d_stem_tma_kernel(const float* __restrict__ src, float* __restrict__ dst, const S2dGeom G) {
    extern __shared__ __align__(128) float s2d_smem[];
    __shared__ __align__(8) uint64_t full[kS2dStages];
    const int wlen = G.W * C, xlen = G.CP * G.Ws, shift = G.pad * C, cp4 = G.CP >> 2;
    const int in_len = BWD ? xlen : 2 * wlen, out_len = BWD ? 2 * wlen : xlen;
    const int stage_len = in_len + out_len;                         // floats; both parts are 16-byte multiples
    const long long rows = (long long)G.N * G.Hs;
    const long long mine = rows > blockIdx.x ? (rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto issue = [&](long long t) {                                 // thread 0: sources of this CTA's t-th row
        const long long ni = blockIdx.x + t * gridDim.x;
        const int st = (int)(t % kS2dStages);
        float* in = s2d_smem + (size_t)st * stage_len;
        const int i = (int)(ni % G.Hs);
        const long long n = ni / G.Hs;
        if (BWD) {
            mbar_arrive_expect_tx(&full[st], (uint32_t)xlen * 4u);
            bulk_g2s(in, src + ni * (long long)xlen, (uint32_t)xlen * 4u, &full[st]);
        } else {
            const int h0 = 2 * i - G.pad;
            const bool v0 = h0 >= 0 && h0 < G.H, v1 = h0 + 1 >= 0 && h0 + 1 < G.H;
            mbar_arrive_expect_tx(&full[st], (uint32_t)((v0 ? wlen : 0) + (v1 ? wlen : 0)) * 4u);
            if (v0) bulk_g2s(in, src + (n * G.H + h0) * (long long)wlen, (uint32_t)wlen * 4u, &full[st]);
            if (v1) bulk_g2s(in + wlen, src + (n * G.H + h0 + 1) * (long long)wlen, (uint32_t)wlen * 4u, &full[st]);
        }
    };

    if (threadIdx.x == 0) {
        for (int k = 0; k < kS2dStages; ++k) mbar_init(&full[k], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (long long t = 0; t < mine && t < kS2dStages; ++t) issue(t);

    for (long long t = 0; t < mine; ++t) {
        const int st = (int)(t % kS2dStages);
        const uint32_t phase = (uint32_t)((t / kS2dStages) & 1);
        float* in = s2d_smem + (size_t)st * stage_len;
        float* out = in + in_len;
        const long long ni = blockIdx.x + t * gridDim.x;
        const int i = (int)(ni % G.Hs);
        const long long n = ni / G.Hs;
        const int h0 = 2 * i - G.pad;
        const bool v0 = h0 >= 0 && h0 < G.H, v1 = h0 + 1 >= 0 && h0 + 1 < G.H;
        // the bulk store that last left from this stage's `out` (kS2dStages rows ago) must have read it completely
        if (threadIdx.x == 0) bulk_wait_read<kS2dStages - 1>();
        __syncthreads();
        mbar_wait(&full[st], phase);
        // One thread per xs pixel j: its 4C floats are 2C consecutive floats of image row 2i-pad followed by the same
        // 2C columns of row 2i+1-pad (columns 2C*j - pad*C ...), so neither direction needs a division.
        if (!BWD) {
            for (int j = threadIdx.x; j < G.Ws; j += kS2dThreads) {
                const int m0 = 2 * C * j - shift;
                float v[4 * C];
#pragma unroll
                for (int k = 0; k < 2 * C; ++k) {
                    const bool ok = (unsigned)(m0 + k) < (unsigned)wlen;
                    v[k] = (ok && v0) ? in[m0 + k] : 0.0f;
                    v[2 * C + k] = (ok && v1) ? in[wlen + m0 + k] : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < C; ++u)
                    reinterpret_cast<float4*>(out)[cp4 * j + u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                for (int u = C; u < cp4; ++u) reinterpret_cast<float4*>(out)[cp4 * j + u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int j = threadIdx.x; j < G.Ws; j += kS2dThreads) {
                const int m0 = 2 * C * j - shift;
                float v[4 * C];
#pragma unroll
                for (int u = 0; u < C; ++u) {
                    const float4 t = reinterpret_cast<const float4*>(in)[cp4 * j + u];
                    v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
                }
#pragma unroll
                for (int k = 0; k < 2 * C; ++k) {
                    if ((unsigned)(m0 + k) < (unsigned)wlen) {
                        out[m0 + k] = v[k];
                        out[wlen + m0 + k] = v[2 * C + k];
                    }
                }
            }
        }
        fence_async_smem();                                          // our STS before the async proxy reads `out`
        __syncthreads();
        if (threadIdx.x == 0) {
            if (!BWD) {
                bulk_s2g(dst + ni * (long long)xlen, out, (uint32_t)xlen * 4u);
            } else {
                if (v0) bulk_s2g(dst + (n * G.H + h0) * (long long)wlen, out, (uint32_t)wlen * 4u);
                if (v1) bulk_s2g(dst + (n * G.H + h0 + 1) * (long long)wlen, out + wlen, (uint32_t)wlen * 4u);
            }
            bulk_commit();
            if (t + kS2dStages < mine) issue(t + kS2dStages);        // every thread is done reading `in` (barrier above)
        }
    }
    if (threadIdx.x == 0) bulk_wait<0>();                            // stores complete before the CTA's memory goes away
}

template <int C, bool BWD>
static bool launch_tma_c(const float* src, float* dst, const S2dGeom& G, cudaStream_t st) {
    const int wlen = G.W * C, xlen = G.CP * G.Ws;
    const size_t smem = (size_t)kS2dStages * (2 * wlen + xlen) * sizeof(float);
    if ((wlen & 3) || (G.CP & 3) || !aligned16(src) || !aligned16(dst) || smem > 160 * 1024) return false;
    auto kernel = s2d_stem_tma_kernel<C, BWD>;
    static int per_sm = -1;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        smem_set = smem;
        per_sm = -1;
    }
    if (per_sm < 0) per_sm = resident_ctas(kernel, kS2dThreads, smem);
    const long long rows = (long long)G.N * G.Hs, cap = (long long)kNumSM * per_sm;
    kernel<<<(unsigned)(rows < cap ? rows : cap), kS2dThreads, smem, st>>>(src, dst, G);
    return true;
}

template <bool BWD>
static bool launch_tma(const float* src, float* dst, const S2dGeom& G, cudaStream_t st) {
    switch (G.C) {
        case 1: return launch_tma_c<1, BWD>(src, dst, G, st);
        case 3: return launch_tma_c<3, BWD>(src, dst, G, st);
        case 4: return launch_tma_c<4, BWD>(src, dst, G, st);
        default: return false;
    }
}

template <bool BWD>
static bool launch_rows(const float* src, float* dst, const S2dGeom& G, cudaStream_t st) {
    if (G.CP != 4 * G.C) return false;
    const long long rows = BWD ? (long long)G.N * G.H : (long long)G.N * G.Hs, cap = (long long)kNumSM * 32;
    const unsigned grid = (unsigned)(rows < cap ? rows : cap);
    switch (G.C) {
        case 1: s2d_stem_rows_kernel<1, BWD><<<grid, 256, 0, st>>>(src, dst, G); return true;
        case 3: s2d_stem_rows_kernel<3, BWD><<<grid, 256, 0, st>>>(src, dst, G); return true;
        case 4: s2d_stem_rows_kernel<4, BWD><<<grid, 256, 0, st>>>(src, dst, G); return true;
        default: return false;
    }
}

static int s2d_geom(int N, int H, int W, int C, int pad, int cpad, S2dGeom& G) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || pad < 0 || ((H + 2 * pad) & 1) || ((W + 2 * pad) & 1)) return OODFQ_EINVAL;
    if (cpad == 0) cpad = 4 * C;
    if (cpad < 4 * C) return OODFQ_EINVAL;
    G.N = N; G.H = H; G.W = W; G.C = C; G.pad = pad; G.CP = cpad;
    G.Hs = (H + 2 * pad) / 2;
    G.Ws = (W + 2 * pad) / 2;
    return OODFQ_OK;
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_s2d_stem_forward(const float* x, float* xs, int N, int H, int W, int C, int pad, int cpad,
                                      oodfq_stream_t stream) {
    if (!x || !xs) return fail(OODFQ_EINVAL, "s2d_stem_forward: null pointer");
    S2dGeom G;
    if (s2d_geom(N, H, W, C, pad, cpad, G) != OODFQ_OK)
        return fail(OODFQ_EINVAL, "s2d_stem_forward: needs a non-empty tensor with even H + 2*pad and W + 2*pad, cpad 0 or >= 4*C");
    const long long total = (long long)N * G.Hs * G.Ws * G.CP;
    if (!launch_tma<false>(x, xs, G, (cudaStream_t)stream) && !launch_rows<false>(x, xs, G, (cudaStream_t)stream)) {
        static const int per_sm = resident_ctas(s2d_stem_kernel<false>, 256);
        long long want = (total + 255) / 256, cap = (long long)kNumSM * per_sm * 4;
        s2d_stem_kernel<false><<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(x, xs, G, total);
    }
    count_launch();
    return check_launch("s2d_stem_forward");
}

extern "C" int oodfq_s2d_stem_backward(const float* grad_xs, float* grad_x, int N, int H, int W, int C, int pad,
                                       int cpad, oodfq_stream_t stream) {
    if (!grad_xs || !grad_x) return fail(OODFQ_EINVAL, "s2d_stem_backward: null pointer");
    S2dGeom G;
    if (s2d_geom(N, H, W, C, pad, cpad, G) != OODFQ_OK)
        return fail(OODFQ_EINVAL, "s2d_stem_backward: needs a non-empty tensor with even H + 2*pad and W + 2*pad, cpad 0 or >= 4*C");
    const long long total = (long long)N * H * W * C;
    if (!launch_tma<true>(grad_xs, grad_x, G, (cudaStream_t)stream) &&
        !launch_rows<true>(grad_xs, grad_x, G, (cudaStream_t)stream)) {
        static const int per_sm = resident_ctas(s2d_stem_kernel<true>, 256);
        long long want = (total + 255) / 256, cap = (long long)kNumSM * per_sm * 4;
        s2d_stem_kernel<true><<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(grad_xs, grad_x, G, total);
    }
    count_launch();
    return check_launch("s2d_stem_backward");
}
