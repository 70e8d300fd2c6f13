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
};

template <bool QUANT, bool XHAT>
__global__ void __launch_bounds__(kBThreads)
bn_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ idx,
                   float* __restrict__ xhat, const PoolGeom G, const BnParams2 P,
                   const float* __restrict__ fq_lo, const float* __restrict__ fq_hi, int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    QParams qp;
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    if ((int)threadIdx.x >= G.lanes_r * G.cols) return;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a[4], b[4], rm[4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { affine2(P, 4 * col + j, a[j], b[j], inv[j]); rm[j] = __ldg(P.rm + 4 * col + j); }
    const long long outs = (long long)G.N * G.Ho * G.Wo;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long o = (long long)blockIdx.x * G.lanes_r + rsub; o < outs; o += (long long)gridDim.x * G.lanes_r) {
        const int wo = (int)(o % G.Wo);
        const int ho = (int)((o / G.Wo) % G.Ho);
        const long long n = o / ((long long)G.Wo * G.Ho);
        float4 v[9];
        bool ok[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int h = 2 * ho - 1 + i, w = 2 * wo - 1 + j;
                ok[i * 3 + j] = (h >= 0) && (h < G.H) && (w >= 0) && (w < G.W);
                if (ok[i * 3 + j]) v[i * 3 + j] = __ldg(x4 + ((n * G.H + h) * G.W + w) * G.cols + col);
            }
        }
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, bx[4] = {0.f, 0.f, 0.f, 0.f};
        int code[4] = {0, 0, 0, 0};
#pragma unroll
        for (int li = 0; li < 9; ++li) {
            if (ok[li]) {
                const float xs[4] = {v[li].x, v[li].y, v[li].z, v[li].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float z = fmaf(xs[j], a[j], b[j]);
                    z = (z != z) ? z : fmaxf(z, 0.0f);
                    const float y = QUANT ? fake_quant_lut(z, qp, lut, qh, qmask) : z;
                    if (y > best[j] || y != y) {          // first maximum in scan order; NaN wins (as max_pool2d)
                        best[j] = y;
                        bx[j] = xs[j];
                        code[j] = li | (z > 0.0f ? 128 : 0);
                    }
                }
            }
        }
        st_out(reinterpret_cast<float4*>(out) + o * G.cols + col, make_float4(best[0], best[1], best[2], best[3]));
        reinterpret_cast<uchar4*>(idx)[o * G.cols + col] =
            make_uchar4((unsigned char)code[0], (unsigned char)code[1], (unsigned char)code[2], (unsigned char)code[3]);
        if (XHAT)
            st_out(reinterpret_cast<float4*>(xhat) + o * G.cols + col,
                   make_float4((bx[0] - rm[0]) * inv[0], (bx[1] - rm[1]) * inv[1], (bx[2] - rm[2]) * inv[2],
                               (bx[3] - rm[3]) * inv[3]));
    }
}

template <bool REDUCE>
__global__ void __launch_bounds__(kBThreads)
bn_pool_bwd_kernel(const float* __restrict__ gout, const uint8_t* __restrict__ idx, const float* __restrict__ xhat,
                   float* __restrict__ gx, const PoolGeom G, const BnParams2 P, Workspace* ws) {
    __shared__ float red[REDUCE ? 2 * kBThreads * 4 : 1];
    const bool active = (int)threadIdx.x < G.lanes_r * G.cols;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f}, sw[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float b, inv; affine2(P, 4 * col + j, a[j], b, inv); }
    }
    const long long pixels = (long long)G.N * G.H * G.W;
    const float4* g4 = reinterpret_cast<const float4*>(gout);
    const uchar4* i4 = reinterpret_cast<const uchar4*>(idx);
    if (active) {
        for (long long p = (long long)blockIdx.x * G.lanes_r + rsub; p < pixels; p += (long long)gridDim.x * G.lanes_r) {
            const int w = (int)(p % G.W);
            const int h = (int)((p / G.W) % G.H);
            const long long n = p / ((long long)G.W * G.H);
            // windows covering row h: ho in [ceil((h-1)/2), floor((h+1)/2)], same for columns
            const int ho0 = h >> 1, ho1 = (h + 1) >> 1, wo0 = w >> 1, wo1 = (w + 1) >> 1;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ih = 0; ih < 2; ++ih) {
                const int ho = ih ? ho1 : ho0;
                if ((ih && ho1 == ho0) || ho >= G.Ho) continue;
#pragma unroll
                for (int iw = 0; iw < 2; ++iw) {
                    const int wo = iw ? wo1 : wo0;
                    if ((iw && wo1 == wo0) || wo >= G.Wo) continue;
                    const int li = (h - 2 * ho + 1) * 3 + (w - 2 * wo + 1);
                    const long long o = ((n * G.Ho + ho) * G.Wo + wo) * G.cols + col;
                    const uchar4 c = __ldg(i4 + o);
                    const float4 g = __ldg(g4 + o);
                    const unsigned char cs[4] = {c.x, c.y, c.z, c.w};
                    const float gs[4] = {g.x, g.y, g.z, g.w};
                    float xh[4] = {0.f, 0.f, 0.f, 0.f};
                    if (REDUCE && li == 4) {                      // this thread sits on the window's centre
                        const float4 t = __ldg(reinterpret_cast<const float4*>(xhat) + o);
                        xh[0] = t.x; xh[1] = t.y; xh[2] = t.z; xh[3] = t.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool on = (cs[j] & 128) != 0;
                        if (on && (cs[j] & 15) == li) acc[j] += gs[j];
                        if (REDUCE && li == 4 && on) { sb[j] += gs[j]; sw[j] = fmaf(gs[j], xh[j], sw[j]); }
                    }
                }
            }
            st_out(reinterpret_cast<float4*>(gx) + p * G.cols + col,
                   make_float4(acc[0] * a[0], acc[1] * a[1], acc[2] * a[2], acc[3] * a[3]));
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
                double* q = ws->bn_partial + ((size_t)blockIdx.x * G.C + 4 * col + j) * 2;
                q[0] = (double)tw;      // dW: xhat already carries 1/sqrt(var+eps)
                q[1] = (double)tb;      // dB
            }
        }
    }
}

static int make_pool_geom(int N, int C, int H, int W, PoolGeom& G) {
    if (C % 4 != 0 || C / 4 > kBThreads) return OODFQ_EINVAL;
    G.N = N; G.C = C; G.H = H; G.W = W;
    G.Ho = (H - 1) / 2 + 1;
    G.Wo = (W - 1) / 2 + 1;
    G.cols = C / 4;
    G.lanes_r = kBThreads / G.cols;
    return OODFQ_OK;
}

}  // namespace oodfq

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
    static const int per_sm = resident_ctas(bn_pool_fwd_kernel<true, true>, kBThreads);
    const long long outs = (long long)N * G.Ho * G.Wo;
    long long want = (outs + G.lanes_r - 1) / G.lanes_r, cap = (long long)kNumSM * per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (quant && xhat) bn_pool_fwd_kernel<true, true><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else if (quant) bn_pool_fwd_kernel<true, false><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else if (xhat) bn_pool_fwd_kernel<false, true><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    else bn_pool_fwd_kernel<false, false><<<grid, kBThreads, 0, st>>>(x, out, idx, xhat, G, P, fq_lo, fq_hi, fq_k);
    count_launch();
    return check_launch("bn_pool_forward");
}

extern "C" int oodfq_bn_pool_backward(const float* grad_out, const uint8_t* idx, const float* xhat, float* grad_x,
                                      int N, int C, int H, int W, const float* weight, const float* bias,
                                      const float* running_mean, const float* running_var, float eps,
                                      double* dwdb, void* workspace, oodfq_stream_t stream) {
    if (!grad_out || !idx || !grad_x || !running_mean || !running_var) return fail(OODFQ_EINVAL, "bn_pool_backward: null pointer");
    if (dwdb && (!xhat || !workspace)) return fail(OODFQ_EINVAL, "bn_pool_backward: parameter gradients need xhat and the workspace");
    PoolGeom G;
    if (make_pool_geom(N, C, H, W, G) != OODFQ_OK || !aligned16(grad_out) || !aligned16(grad_x) || (xhat && !aligned16(xhat)))
        return fail(OODFQ_EINVAL, "bn_pool_backward: needs C %% 4 == 0, C <= 1024 and aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const BnParams2 P{weight, bias, running_mean, running_var, eps};
    static const int occ[2] = {resident_ctas(bn_pool_bwd_kernel<false>, kBThreads), resident_ctas(bn_pool_bwd_kernel<true>, kBThreads)};
    const long long pixels = (long long)N * H * W;
    long long want = (pixels + G.lanes_r - 1) / G.lanes_r, cap = (long long)kNumSM * occ[dwdb ? 1 : 0];
    const long long table = (long long)kMaxBnSplit * kMaxBnChannels / C;
    if (dwdb && cap > table) cap = table;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (dwdb) bn_pool_bwd_kernel<true><<<grid, kBThreads, 0, st>>>(grad_out, idx, xhat, grad_x, G, P, ws);
    else bn_pool_bwd_kernel<false><<<grid, kBThreads, 0, st>>>(grad_out, idx, xhat, grad_x, G, P, ws);
    count_launch();
    int rc = check_launch("bn_pool_backward");
    if (rc != OODFQ_OK || !dwdb) return rc;
    bn_nhwc_fold_kernel<<<(C + kBThreads / 32 - 1) / (kBThreads / 32), kBThreads, 0, st>>>(ws->bn_partial, C, (int)grid, dwdb);
    count_launch();
    return check_launch("bn_pool_backward(fold)");
}
