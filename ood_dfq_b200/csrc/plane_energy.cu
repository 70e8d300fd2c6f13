// Per-(image, channel) mean of squares of an activation tensor, and its backward.
//
// This is the inner reduction of the reference's feature-alignment loss: Trainer.channel_attention
// (trainer_direct.py:382-383) computes  F.normalize(x.pow(2).mean([2,3]))  on a clone of every residual
// body output of student and teacher (hooks :432-440, loss :325-330).  In eager PyTorch that is a clone,
// a pow and a mean forward (20 B/elem) and a long strided tape backward (~24 B/elem); after the BatchNorm
// fusion it was the largest remaining non-convolution cost of the step (profiles/r1_step_share_fused_channels_last.txt:
// the two `elementwise_kernel<128,2>` entries).  Here:
//
//   forward   E[n,c] = mean_{hw} x[n,c,hw]^2                 one read            (4 B/elem)
//   backward  grad_x[n,c,hw] = gE[n,c] * 2/HW * x[n,c,hw]    one read, one write (8 B/elem)
//
// NCHW: one warp per (n,c) plane.  channels_last: one CTA per (image, row chunk), a thread owns a 128-bit
// column (4 channels) of that image's rows; chunk partials are folded by a second tiny launch (ordered sums,
// no atomics).  Roofline: HBM.
#include "bn_geom.cuh"

namespace oodfq {

constexpr int kEWarps = kBThreads / 32;

__global__ void __launch_bounds__(kBThreads)
energy_nchw_fwd_kernel(const float* __restrict__ x, float* __restrict__ e, long long planes, int HW, float inv_hw,
                       int vec) {
    const long long p = (long long)blockIdx.x * kEWarps + (threadIdx.x >> 5);
    if (p >= planes) return;
    const int lane = threadIdx.x & 31;
    const float* px = x + p * HW;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(px);
        const int n4 = HW >> 2;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
            float4 v = ld_stream(p4 + i);
            s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1); s2 = fmaf(v.z, v.z, s2); s3 = fmaf(v.w, v.w, s3);
        }
    } else {
#pragma unroll 4
        for (int i = lane; i < HW; i += 32) { float v = ld_stream(px + i); s0 = fmaf(v, v, s0); }
    }
    float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) e[p] = s * inv_hw;
}

__global__ void __launch_bounds__(kBThreads)
energy_nchw_bwd_kernel(const float* __restrict__ x, const float* __restrict__ ge, float* __restrict__ gx,
                       long long planes, int HW, float two_inv_hw, int vec) {
    const long long p = (long long)blockIdx.x * kEWarps + (threadIdx.x >> 5);
    if (p >= planes) return;
    const int lane = threadIdx.x & 31;
    const float c = __ldg(ge + p) * two_inv_hw;
    const float* px = x + p * HW;
    float* pg = gx + p * HW;
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(px);
        float4* g4 = reinterpret_cast<float4*>(pg);
        const int n4 = HW >> 2;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
            float4 v = ld_stream(p4 + i);
            st_out(g4 + i, make_float4(c * v.x, c * v.y, c * v.z, c * v.w));
        }
    } else {
#pragma unroll 4
        for (int i = lane; i < HW; i += 32) pg[i] = c * ld_stream(px + i);
    }
}

// channels_last: image n is the contiguous block of HW rows x C channels starting at row n*HW
struct EnergyNhwc {
    int N, C, HW, cols, lanes_r, col_blocks, chunks, rows_per_chunk;
};

template <bool BWD>
__global__ void __launch_bounds__(kBThreads)
energy_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ ge, float* __restrict__ out,
                   const EnergyNhwc G, float scale) {
    __shared__ float red[BWD ? 1 : kBThreads * 4];
    const int n = blockIdx.x / G.chunks, ck = blockIdx.x % G.chunks;
    const int wcols = G.cols < kBThreads ? G.cols : kBThreads;
    const bool active = (int)threadIdx.x < G.lanes_r * wcols;
    const int rsub = threadIdx.x / wcols;
    const int r_begin = ck * G.rows_per_chunk;
    const int r_end = min(G.HW, r_begin + G.rows_per_chunk);
    const float4* xi = reinterpret_cast<const float4*>(x) + (long long)n * G.HW * G.cols;
    for (int cb = 0; cb < G.col_blocks; ++cb) {
        const int col = cb * kBThreads + threadIdx.x % wcols;
        const bool on = active && col < G.cols;
        float s[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f};
        if (BWD && on) {
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j] = __ldg(ge + (long long)n * G.C + 4 * col + j) * scale;
        }
        if (on) {
            for (int r = r_begin + rsub; r < r_end; r += kDepth * G.lanes_r) {
                float4 v[kDepth];
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const int rr = r + d * G.lanes_r;
                    if (rr < r_end) v[d] = ld_stream(xi + (long long)rr * G.cols + col);
                }
#pragma unroll
                for (int d = 0; d < kDepth; ++d) {
                    const int rr = r + d * G.lanes_r;
                    if (rr < r_end) {
                        if (BWD) {
                            st_out(reinterpret_cast<float4*>(out) + ((long long)n * G.HW + rr) * G.cols + col,
                                   make_float4(c[0] * v[d].x, c[1] * v[d].y, c[2] * v[d].z, c[3] * v[d].w));
                        } else {
                            s[0] = fmaf(v[d].x, v[d].x, s[0]); s[1] = fmaf(v[d].y, v[d].y, s[1]);
                            s[2] = fmaf(v[d].z, v[d].z, s[2]); s[3] = fmaf(v[d].w, v[d].w, s[3]);
                        }
                    }
                }
            }
        }
        if (!BWD) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = on ? s[j] : 0.f;
            __syncthreads();
            if (on && rsub == 0) {
                const int lc = threadIdx.x % wcols;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = 0.f;
                    for (int l = 0; l < G.lanes_r; ++l) t += red[(l * wcols + lc) * 4 + j];
                    // partial[chunk][n][c]
                    out[((long long)ck * G.N + n) * G.C + 4 * col + j] = t * scale;
                }
            }
        }
    }
}

__global__ void energy_fold_kernel(const float* __restrict__ partial, float* __restrict__ e, long long nc, int chunks) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    float t = 0.f;
    for (int k = 0; k < chunks; ++k) t += partial[(long long)k * nc + i];
    e[i] = t;
}

static EnergyNhwc make_energy_nhwc(int N, int C, int HW, int slots) {
    EnergyNhwc G;
    G.N = N; G.C = C; G.HW = HW; G.cols = C / 4;
    G.lanes_r = G.cols <= kBThreads ? kBThreads / G.cols : 1;
    G.col_blocks = (G.cols + kBThreads - 1) / kBThreads;
    int passes = (HW + G.lanes_r * kDepth - 1) / (G.lanes_r * kDepth);     // pipeline rounds per image
    int want = slots / N;                          // CTAs per image: fill the resident slots once, never 1.x waves
    if (want > passes) want = passes;
    if (want > 16) want = 16;
    if (want < 1) want = 1;
    G.chunks = want;
    G.rows_per_chunk = (HW + want - 1) / want;
    return G;
}

}  // namespace oodfq

using namespace oodfq;

extern "C" size_t oodfq_channel_energy_scratch_floats(int N, int C) { return (size_t)16 * (size_t)N * (size_t)C; }

extern "C" int oodfq_channel_energy_forward(const float* x, float* e, int N, int C, long long HW, int flags,
                                            float* scratch, oodfq_stream_t stream) {
    if (!x || !e) return fail(OODFQ_EINVAL, "channel_energy_forward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0 || HW > 0x7fffffffLL) return fail(OODFQ_EINVAL, "channel_energy_forward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const float inv = (float)(1.0 / (double)HW);
    if (flags & OODFQ_BN_NHWC) {
        if ((C % 4) != 0 || !aligned16(x)) return fail(OODFQ_EINVAL, "channel_energy_forward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        if (!scratch) return fail(OODFQ_EINVAL, "channel_energy_forward: NHWC needs the scratch buffer");
        static const int per_sm = resident_ctas(energy_nhwc_kernel<false>, kBThreads);
        const EnergyNhwc G = make_energy_nhwc(N, C, (int)HW, kNumSM * per_sm);
        energy_nhwc_kernel<false><<<(unsigned)N * G.chunks, kBThreads, 0, st>>>(x, nullptr, G.chunks == 1 ? e : scratch, G, inv);
        count_launch();
        int rc = check_launch("channel_energy_forward");
        if (rc != OODFQ_OK || G.chunks == 1) return rc;
        const long long nc = (long long)N * C;
        energy_fold_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(scratch, e, nc, G.chunks);
        count_launch();
        return check_launch("channel_energy_forward(fold)");
    }
    const long long planes = (long long)N * C;
    const int vec = (HW % 4 == 0) && aligned16(x);
    energy_nchw_fwd_kernel<<<(unsigned)((planes + kEWarps - 1) / kEWarps), kBThreads, 0, st>>>(x, e, planes, (int)HW, inv, vec);
    count_launch();
    return check_launch("channel_energy_forward");
}

extern "C" int oodfq_channel_energy_backward(const float* x, const float* grad_e, float* grad_x, int N, int C,
                                             long long HW, int flags, oodfq_stream_t stream) {
    if (!x || !grad_e || !grad_x) return fail(OODFQ_EINVAL, "channel_energy_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0 || HW > 0x7fffffffLL) return fail(OODFQ_EINVAL, "channel_energy_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const float two_inv = (float)(2.0 / (double)HW);
    if (flags & OODFQ_BN_NHWC) {
        if ((C % 4) != 0 || !aligned16(x) || !aligned16(grad_x))
            return fail(OODFQ_EINVAL, "channel_energy_backward: NHWC needs C %% 4 == 0 and 16-byte alignment");
        static const int per_sm = resident_ctas(energy_nhwc_kernel<true>, kBThreads);
        const EnergyNhwc G = make_energy_nhwc(N, C, (int)HW, kNumSM * per_sm);
        energy_nhwc_kernel<true><<<(unsigned)N * G.chunks, kBThreads, 0, st>>>(x, grad_e, grad_x, G, two_inv);
        count_launch();
        return check_launch("channel_energy_backward");
    }
    const long long planes = (long long)N * C;
    const int vec = (HW % 4 == 0) && aligned16(x) && aligned16(grad_x);
    energy_nchw_bwd_kernel<<<(unsigned)((planes + kEWarps - 1) / kEWarps), kBThreads, 0, st>>>(x, grad_e, grad_x, planes, (int)HW, two_inv, vec);
    count_launch();
    return check_launch("channel_energy_backward");
}
