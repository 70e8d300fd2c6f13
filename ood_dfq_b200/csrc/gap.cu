// Global average pool behind the last residual unit, channels_last: [N, H*W, C] -> [N, C], and its backward.
//
// The carrier networks end in  features.final_pool = AvgPool2d(7 | 8)  over a plane of exactly that size
// (pytorchcv ResNet behind ptcv_get_model, main_direct.py:380-397) -- the tensor between the last
// Sequential(ReLU, QuantAct) and Quant_Linear.  ATen's generic channels_last pooling kernels decode
// (n, h, w, c) per element with integer divisions and walk the covering windows: the backward of this 25.7 MB
// tensor takes 90 us per launch, four launches per QAT step (1.1 % of the 224x224 step,
// profiles/r2_step_share_top120.txt), the forward 14 us.  Here:
//
//   forward   y[n,c]     = (sum_{hw, in order} x[n,hw,c]) / HW      one read             (4 B/elem)
//   backward  gx[n,hw,c] = gy[n,c] / HW                             one write            (4 B/elem)
//
// Same arithmetic as avg_pool2d_out_cuda_frame_nhwc / avg_pool2d_backward_out_cuda_frame_nhwc for a single
// window (fp32 running sum in row-major window order, one IEEE division by the window size), so results are
// bit-identical to nn.AvgPool2d.  Roofline: HBM (launch-latency-bound at these sizes).
#include "bn_geom.cuh"

namespace oodfq {

constexpr int kGapDepth = 8;      // rows in flight per thread

// a thread owns one 128-bit column (4 channels) of one image and adds its rows in order
__global__ void __launch_bounds__(kBThreads)
gap_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long ncols, int cols, int HW, float hw) {
    const long long t = (long long)blockIdx.x * kBThreads + threadIdx.x;       // = n * cols + col
    if (t >= ncols) return;
    const long long n = t / cols;
    const int col = (int)(t % cols);
    const float4* p = reinterpret_cast<const float4*>(x) + n * HW * cols + col;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < HW; r += kGapDepth) {
        float4 v[kGapDepth];
#pragma unroll
        for (int d = 0; d < kGapDepth; ++d)
            if (r + d < HW) v[d] = ld_stream(p + (long long)(r + d) * cols);
#pragma unroll
        for (int d = 0; d < kGapDepth; ++d)
            if (r + d < HW) {
                s[0] = __fadd_rn(s[0], v[d].x); s[1] = __fadd_rn(s[1], v[d].y);
                s[2] = __fadd_rn(s[2], v[d].z); s[3] = __fadd_rn(s[3], v[d].w);
            }
    }
    reinterpret_cast<float4*>(y)[t] = make_float4(__fdiv_rn(s[0], hw), __fdiv_rn(s[1], hw), __fdiv_rn(s[2], hw), __fdiv_rn(s[3], hw));
}

// a thread owns one column of one image, divides once and writes that image's rows
__global__ void __launch_bounds__(kBThreads)
gap_nhwc_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long ncols, int cols, int HW, int split,
                    float hw) {
    const long long t = ((long long)blockIdx.x / split) * kBThreads + threadIdx.x;
    if (t >= ncols) return;
    const int part = blockIdx.x % split;                                       // this CTA's share of the rows
    const long long n = t / cols;
    const int col = (int)(t % cols);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gy) + t);
    // ATen accumulates `gradient += g / divide_factor` from zero: the sum with +0 turns a -0 quotient into +0
    const float4 o = make_float4(__fadd_rn(0.0f, __fdiv_rn(g.x, hw)), __fadd_rn(0.0f, __fdiv_rn(g.y, hw)),
                                 __fadd_rn(0.0f, __fdiv_rn(g.z, hw)), __fadd_rn(0.0f, __fdiv_rn(g.w, hw)));
    float4* p = reinterpret_cast<float4*>(gx) + n * HW * cols + col;
    for (int r = part; r < HW; r += split) st_out(p + (long long)r * cols, o);
}

}  // namespace oodfq

using namespace oodfq;

static int gap_check(const char* what, const void* a, const void* b, int N, int C, long long HW, int flags) {
    if (!a || !b) return fail(OODFQ_EINVAL, "%s: null pointer", what);
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "%s: empty tensor", what);
    if (!(flags & OODFQ_BN_NHWC)) return fail(OODFQ_EINVAL, "%s: channels_last only", what);
    if (C % 4 != 0 || HW > (1 << 24) || !aligned16(a) || !aligned16(b))
        return fail(OODFQ_EINVAL, "%s: needs C %% 4 == 0, H*W <= 2^24 and 16-byte aligned buffers", what);
    return OODFQ_OK;
}

extern "C" int oodfq_global_avgpool_forward(const float* x, float* y, int N, int C, long long HW, int flags,
                                            oodfq_stream_t stream) {
    int rc = gap_check("global_avgpool_forward", x, y, N, C, HW, flags);
    if (rc != OODFQ_OK) return rc;
    const int cols = C / 4;
    const long long ncols = (long long)N * cols;
    gap_nhwc_fwd_kernel<<<(unsigned)((ncols + kBThreads - 1) / kBThreads), kBThreads, 0, (cudaStream_t)stream>>>(
        x, y, ncols, cols, (int)HW, (float)HW);
    count_launch();
    return check_launch("global_avgpool_forward");
}

extern "C" int oodfq_global_avgpool_backward(const float* grad_y, float* grad_x, int N, int C, long long HW, int flags,
                                             oodfq_stream_t stream) {
    int rc = gap_check("global_avgpool_backward", grad_y, grad_x, N, C, HW, flags);
    if (rc != OODFQ_OK) return rc;
    const int cols = C / 4;
    const long long ncols = (long long)N * cols;
    const long long base = (ncols + kBThreads - 1) / kBThreads;
    // enough CTAs to fill the machine: the rows of an image are dealt out to `split` CTAs
    long long split = (4LL * kNumSM + base - 1) / base;
    if (split > HW) split = HW;
    if (split < 1) split = 1;
    gap_nhwc_bwd_kernel<<<(unsigned)(base * split), kBThreads, 0, (cudaStream_t)stream>>>(grad_y, grad_x, ncols, cols,
                                                                                        (int)HW, (int)split, (float)HW);
    count_launch();
    return check_launch("global_avgpool_backward");
}
