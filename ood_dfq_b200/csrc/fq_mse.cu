// QuantAct_MSE range search: all clip candidates scored in ONE pass over the activation tensor.
//
// Replaces the loop of QuantAct_MSE.forward (quantization_utils/quant_modules.py:160-178): for i in 0..79 the
// reference shrinks the data range by (1 - 0.01 i), fake-quantises a clone of x with it (find_MSESmallest,
// quant_utils.py:36-47: six ATen passes) and scores mean(|x - fq(x)|^2.4) (lp_loss, quant_utils.py:26-33: three
// more passes and a host sync for `score < best_score`) -- ~80 x 72 B/elem and 80 syncs.  Here x is read once
// (4 B/elem); the 80 fake-quantisations of an element happen in registers, so the kernel is bound by the SFU /
// issue rate (80 x [5-op code, table look-up, |d|^2.4 via lg2/ex2]), not by HBM: ~1.2 k instructions per element.
//
//   mse_scores_kernel   a CTA stages tiles of x in shared memory; warp w scores candidates {w, w+8, ...} over the
//                       whole tile with register accumulators (candidate parameters and dequantisation tables
//                       live in shared memory), so every element is visited by all 8 warps but each (element,
//                       candidate) pair exactly once.  One fp64 partial per (CTA, candidate).
//   mse_select_kernel   folds the partials in CTA order, takes the first strict minimum exactly like the
//                       reference's `if score < best_score` scan (best_score starts at 1e10), then the plain EMA
//                       of quant_modules.py:176-178 (no bias correction), all on the device: no host sync.
//
// Arithmetic: candidate range = fp32(data_min * fp32(1 - 0.01 i)) (a 0-dim tensor times a Python float), the
// quantisation parameters and codes are the reference's fp32 sequence (common.cuh); the score itself is a
// reduction whose order differs from ATen's, so it agrees to rounding, and the selected candidate agrees
// whenever two candidates are not tied to within that rounding.
#include <cmath>

#include "common.cuh"

namespace oodfq {

constexpr int kMseThreads = 256;
constexpr int kMseWarps = kMseThreads / 32;
constexpr int kMseTile = 2048;           // floats per staged tile
constexpr int kMseMaxCand = 96;
constexpr int kMseLutBits = 5;           // tables for k <= 5 (96 x 32 floats = 12 KB); true division above
constexpr int kMsePerWarp = kMseMaxCand / kMseWarps;

__device__ __forceinline__ float pow_abs(float d, float p) {
    // |d|^p = 2^(p * log2|d|); log2(0) = -inf -> 0
    return exp2f(p * __log2f(fabsf(d)));
}

template <bool LUT>
__global__ void __launch_bounds__(kMseThreads)
mse_scores_kernel(const float* __restrict__ x, long long numel, const float* __restrict__ data_mm, int k, int ncand,
                  double step, float p, double* __restrict__ partial /* [grid][ncand] */) {
    __shared__ float tile[kMseTile];
    __shared__ float s_scale[kMseMaxCand], s_zp[kMseMaxCand];
    __shared__ float s_lut[LUT ? kMseMaxCand * (1 << kMseLutBits) : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float qlo = -(float)(1 << (k - 1)), qhi = (float)(1 << (k - 1)) - 1.0f;
    const int h = 1 << (k - 1), mask = (1 << k) - 1, lut_n = 1 << k;
    for (int c = threadIdx.x; c < ncand; c += kMseThreads) {
        const float f = (float)(1.0 - (double)c * step);              // Python: 1.0 - (i * 0.01), then fp32
        const QParams q = make_qparams(__fmul_rn(__ldg(data_mm), f), __fmul_rn(__ldg(data_mm + 1), f), k);
        s_scale[c] = q.scale;
        s_zp[c] = q.zp;
    }
    __syncthreads();
    if (LUT) {
        for (int e = threadIdx.x; e < ncand * lut_n; e += kMseThreads) {
            const int c = e / lut_n, j = e % lut_n;
            const QParams q = given_qparams(s_scale[c], s_zp[c], k);
            s_lut[c * (1 << kMseLutBits) + j] = value_of<false>(__fadd_rn((float)j, qlo), q);
        }
    }
    float acc[kMsePerWarp];
#pragma unroll
    for (int i = 0; i < kMsePerWarp; ++i) acc[i] = 0.0f;
    const long long tiles = (numel + kMseTile - 1) / kMseTile;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        __syncthreads();                                   // previous tile fully consumed (and tables ready)
        const long long base = t * kMseTile;
        const int len = (int)((numel - base) < kMseTile ? (numel - base) : kMseTile);
        for (int e = threadIdx.x; e < kMseTile; e += kMseThreads) tile[e] = e < len ? ld_stream(x + base + e) : 0.0f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kMsePerWarp; ++i) {
            const int c = warp + i * kMseWarps;
            if (c < ncand) {
                const float scale = s_scale[c], zp = s_zp[c];
                const QParams q = given_qparams(scale, zp, k);
                float a = 0.0f;
                for (int e = lane; e < len; e += 32) {
                    const float v = tile[e];
                    const float code = code_of<false>(v, q);
                    float y;
                    if (LUT) {
                        y = s_lut[c * (1 << kMseLutBits) + lut_index(code, h, mask)];
                        y = (code != code) ? code : y;
                    } else {
                        y = value_of<false>(code, q);
                    }
                    a += pow_abs(__fsub_rn(v, y), p);
                }
                acc[i] += a;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kMsePerWarp; ++i) {
        const int c = warp + i * kMseWarps;
        const float s = warp_sum(acc[i]);
        if (lane == 0 && c < ncand) partial[(long long)blockIdx.x * ncand + c] = (double)s;
    }
}

__global__ void mse_select_kernel(const double* __restrict__ partial, int nparts, int ncand, long long numel,
                                  const float* __restrict__ data_mm, double step, float* x_min, float* x_max,
                                  const float* beta, float* beta_t, float* cur_min, float* cur_max,
                                  float* scores /* nullable [ncand] */, int* chosen /* nullable */) {
    __shared__ float s_score[kMseMaxCand];
    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        double t = 0.0;
        for (int b = 0; b < nparts; ++b) t += partial[(long long)b * ncand + c];
        s_score[c] = (float)(t / (double)numel);
        if (scores) scores[c] = s_score[c];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float best = 1e+10f;
    int keep = -1;
    for (int c = 0; c < ncand; ++c)
        if (s_score[c] < best) { best = s_score[c]; keep = c; }       // strict: the first minimum wins
    if (chosen) *chosen = keep;
    const float dmin = data_mm[0], dmax = data_mm[1];
    if (cur_min) *cur_min = dmin;                                        // quant_modules.py:150-151
    if (cur_max) *cur_max = dmax;
    if (keep < 0) keep = 0;          // every score NaN / huge: the reference would fail on an unbound name
    const float f = (float)(1.0 - (double)keep * step);
    const float lo = __fmul_rn(dmin, f), hi = __fmul_rn(dmax, f);
    const float b = *beta, omb = __fsub_rn(1.0f, b);
    *beta_t = __fmul_rn(*beta_t, b);                                                          // :176
    *x_min = __fadd_rn(__fmul_rn(*x_min, b), __fmul_rn(lo, omb));                             // :177
    *x_max = __fadd_rn(__fmul_rn(*x_max, b), __fmul_rn(hi, omb));                             // :178
}

}  // namespace oodfq

using namespace oodfq;

extern "C" size_t oodfq_act_mse_scratch_doubles(int ncand) { return (size_t)kNumSM * 4 * (size_t)ncand; }

extern "C" int oodfq_act_mse_search(const float* x, long long numel, const float* data_minmax, int k, int ncand,
                                    double step, float p, float* x_min, float* x_max, const float* beta,
                                    float* beta_t, float* cur_min, float* cur_max, double* scratch, float* scores,
                                    int* chosen, oodfq_stream_t stream) {
    if (!x || !data_minmax || !x_min || !x_max || !beta || !beta_t || !scratch)
        return fail(OODFQ_EINVAL, "act_mse_search: null pointer");
    if (numel <= 0) return fail(OODFQ_EINVAL, "act_mse_search: empty tensor");
    if (k < 1 || k > 16) return fail(OODFQ_EINVAL, "act_mse_search: k=%d outside [1,16]", k);
    if (ncand < 1 || ncand > kMseMaxCand) return fail(OODFQ_EINVAL, "act_mse_search: %d candidates (max %d)", ncand, kMseMaxCand);
    cudaStream_t st = (cudaStream_t)stream;
    const long long tiles = (numel + kMseTile - 1) / kMseTile, cap = (long long)kNumSM * 4;
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (k <= kMseLutBits) mse_scores_kernel<true><<<grid, kMseThreads, 0, st>>>(x, numel, data_minmax, k, ncand, step, p, scratch);
    else mse_scores_kernel<false><<<grid, kMseThreads, 0, st>>>(x, numel, data_minmax, k, ncand, step, p, scratch);
    count_launch();
    int rc = check_launch("act_mse_search(scores)");
    if (rc != OODFQ_OK) return rc;
    mse_select_kernel<<<1, 128, 0, st>>>(scratch, grid, ncand, numel, data_minmax, step, x_min, x_max, beta, beta_t,
                                         cur_min, cur_max, scores, chosen);
    count_launch();
    return check_launch("act_mse_search(select)");
}
