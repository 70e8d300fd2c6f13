// Tail of a residual unit as ONE kernel each way, channels_last:
//
//     y = QuantAct(ReLU( BN1(x1) + id ))        id = r   or   BN2(r)        (+ E[n,c] = mean_hw BN1(x1)^2)
//
// After quantize_model (main_direct.py:444-479) every residual unit of the student ends in
//   body.conv2.bn (eval-mode affine)  ->  + identity (plain, or identity_conv.bn)  ->  Sequential(ReLU, QuantAct)
// and the trainer's feature-alignment hook (trainer_direct.py:432-440, :382-383) reads the body output z1 once
// more for its per-(image, channel) mean of squares.  With the BatchNorm fusion of bn_fused.cu that is still
// four or five kernels forward (BN 8 B/elem, [BN 8], add 12, ReLU+QuantAct 8, energy 4) and, backward, the ReLU
// mask (12), the energy gradient (8), autograd's accumulation add (12) and the BN backward(s) (12 each):
// the adds and masks were 6 ms of the 47 ms ImageNet step (profiles/r1_step_share_pruned_stem2.txt).  Here:
//
//   forward   read x1, r; write y                       12 B/elem   (E comes out of the same read)
//   backward  read gy, x1, r; write gx1, gr             20 B/elem   (+4 with a second output gradient, see below)
//             s = z1 + id;  g = gy * [s > 0]  (ReLU; the quantiser is the identity STE, quant_utils.py:159-161)
//             dz1 = g + gE[n,c] * 2/HW * z1;  gx1 = a1 * dz1;  gr = g  or  a2 * g
//             dW1 = sum dz1 * xhat1, dB1 = sum dz1  (and dW2, dB2 over g) from the same pass
//
// A unit's output feeds two consumers in the next unit (its body and its identity path).  Autograd would sum their
// gradients with one more 12 B/elem kernel per unit and sweep (5.5 % of the step); instead the forward hands out
// two handles of y and the backward takes both gradients (grad_y2, nullable) and adds them in registers.
//
// Every intermediate is rounded exactly where the unfused chain rounds it (affine by FFMA as bn_fused.cu, the add
// and the gradient sum as separate fp32 roundings), so y and the gradients are bit-identical to the chain of
// bn_fused.cu + ATen add + fq_elementwise.cu; only the fp32 summation order of E, dW, dB differs.
//
// Mapping: a CTA owns (image, chunk of rows) work items -- the energy and its gradient coefficient are per
// image -- and a thread one 128-bit column (4 channels) of those rows, as plane_energy.cu.  Parameter-gradient
// partials stay in registers across the items of a CTA and are folded in CTA order (deterministic, no atomics).
// Roofline: HBM.
#include <type_traits>

#include "bn_geom.cuh"

namespace oodfq {

struct TailBn {
    const float* w;
    const float* b;
    const float* rm;
    const float* rv;     // NULL: this BatchNorm is absent (plain identity)
    float eps;
};

__device__ __forceinline__ void tail_affine(const TailBn& P, int c, float& a, float& b, float& invstd) {
    invstd = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(P.rv + c), P.eps)));
    a = __fmul_rn(P.w ? __ldg(P.w + c) : 1.0f, invstd);
    b = __fsub_rn(P.b ? __ldg(P.b + c) : 0.0f, __fmul_rn(__ldg(P.rm + c), a));
}

struct TailGeom {
    int N, C, HW, cols, lanes_r, chunks, rows_per_chunk;
};

constexpr int kTailFwdDepth = 4;    // rows in flight per thread: 8 loads forward,
constexpr int kTailBwdDepth = 2;    // 6 loads backward (more state in registers)

template <bool QUANT, bool IDBN, bool ENERGY>
__global__ void __launch_bounds__(kBThreads)
res_tail_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ r, float* __restrict__ y,
                    float* __restrict__ epart, uint8_t* __restrict__ mask, const TailGeom G, const TailBn P1,
                    const TailBn P2, float inv_hw, const float* __restrict__ fq_lo, const float* __restrict__ fq_hi,
                    int fq_k) {
    __shared__ float lut[QUANT ? kLutMax : 1];
    __shared__ float red[ENERGY ? kBThreads * 4 : 1];
    QParams qp = given_qparams(1.0f, 0.0f, 1);
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    float lowc = 0.0f;
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        lowc = relu_lower_bound(qp);
        build_lut(lut, qp, fq_k, threadIdx.x, kBThreads);
        __syncthreads();
    }
    const bool active = (int)threadIdx.x < G.lanes_r * G.cols;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a1[4], b1[4], a2[4], b2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a1[j] = b1[j] = b2[j] = 0.f; a2[j] = 1.f;
        if (active) {
            float inv;
            tail_affine(P1, 4 * col + j, a1[j], b1[j], inv);
            if (IDBN) tail_affine(P2, 4 * col + j, a2[j], b2[j], inv);
        }
    }
    const int items = G.N * G.chunks;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / G.chunks, ck = item % G.chunks;
        const int r_begin = ck * G.rows_per_chunk, r_end = min(G.HW, r_begin + G.rows_per_chunk);
        const long long base = (long long)n * G.HW * G.cols + col;
        const float4* p1 = reinterpret_cast<const float4*>(x1) + base;
        const float4* p2 = reinterpret_cast<const float4*>(r) + base;
        float4* py = reinterpret_cast<float4*>(y) + base;
        uint8_t* pm = mask ? mask + base : nullptr;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            for (int row = r_begin + rsub; row < r_end; row += kTailFwdDepth * G.lanes_r) {
                float4 u[kTailFwdDepth], v[kTailFwdDepth];
#pragma unroll
                for (int d = 0; d < kTailFwdDepth; ++d) {
                    const int rr = row + d * G.lanes_r;
                    if (rr < r_end) {
                        u[d] = ld_stream(p1 + (long long)rr * G.cols);
                        v[d] = ld_stream(p2 + (long long)rr * G.cols);
                    }
                }
#pragma unroll
                for (int d = 0; d < kTailFwdDepth; ++d) {
                    const int rr = row + d * G.lanes_r;
                    if (rr < r_end) {
                        const float xs[4] = {u[d].x, u[d].y, u[d].z, u[d].w};
                        const float rs[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
                        float o[4];
                        unsigned open = 0;              // bit j: the ReLU lets the gradient of channel j through
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float z1 = fmaf(xs[j], a1[j], b1[j]);
                            if (ENERGY) s[j] = fmaf(z1, z1, s[j]);
                            const float id = IDBN ? fmaf(rs[j], a2[j], b2[j]) : rs[j];
                            float t = __fadd_rn(z1, id);
                            open |= (t <= 0.0f) ? 0u : (1u << j);     // aten::threshold_backward: NaN passes
                            // ReLU (NaN kept) then the quantiser; with QUANT the ReLU is the quantiser's lower clamp
                            o[j] = QUANT ? relu_fake_quant_lut(t, qp, lowc, lut, qh, qmask) : relu_keep_nan(t);
                        }
                        st_out(py + (long long)rr * G.cols, make_float4(o[0], o[1], o[2], o[3]));
                        if (pm) pm[(long long)rr * G.cols] = (uint8_t)open;
                    }
                }
            }
        }
        if (ENERGY) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) red[threadIdx.x * 4 + j] = active ? s[j] : 0.f;
            __syncthreads();
            if (active && rsub == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float t = 0.f;
                    for (int l = 0; l < G.lanes_r; ++l) t += red[(l * G.cols + col) * 4 + j];
                    epart[((long long)ck * G.N + n) * G.C + 4 * col + j] = t * inv_hw;     // partial[chunk][n][c]
                }
            }
        }
    }
}

__global__ void res_tail_fold_energy_kernel(const float* __restrict__ partial, float* __restrict__ e, long long nc,
                                            int chunks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    float t = 0.f;
    for (int k = 0; k < chunks; ++k) t += partial[(long long)k * nc + i];
    e[i] = t;
}

// MASK: the forward left one byte per 128-bit column (bit j = the ReLU of channel j is open), so the identity `r` is
// only read when its BatchNorm's weight gradient needs it (IDBN && REDUCE) and the body output `x1` only when the
// energy gradient or BN1's weight gradient does: 16.25 instead of 20 B/elem for the common case.
template <bool IDBN, bool ENERGY, bool REDUCE, bool MASK>
__global__ void __launch_bounds__(kBThreads)
res_tail_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ gy2, const float* __restrict__ ge,
                    const float* __restrict__ x1, const float* __restrict__ r, const uint8_t* __restrict__ mask,
                    float* __restrict__ gx1, float* __restrict__ gr, const TailGeom G, const TailBn P1,
                    const TailBn P2, float two_inv_hw, double* __restrict__ part) {
    constexpr bool NEED_X1 = ENERGY || REDUCE || !MASK;
    constexpr bool NEED_R = !MASK || (IDBN && REDUCE);
    __shared__ __align__(16) float red[REDUCE ? (IDBN ? 4 : 2) * kBThreads * 4 : 4];
    const bool active = (int)threadIdx.x < G.lanes_r * G.cols;
    const int col = threadIdx.x % G.cols, rsub = threadIdx.x / G.cols;
    float a1[4], b1[4], rm1[4], a2[4], b2[4], rm2[4];
    float sb1[4], sw1[4], sb2[4], sw2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a1[j] = b1[j] = rm1[j] = b2[j] = rm2[j] = 0.f; a2[j] = 1.f;
        sb1[j] = sw1[j] = sb2[j] = sw2[j] = 0.f;
        if (active) {
            float inv;
            tail_affine(P1, 4 * col + j, a1[j], b1[j], inv);
            rm1[j] = __ldg(P1.rm + 4 * col + j);
            if (IDBN) { tail_affine(P2, 4 * col + j, a2[j], b2[j], inv); rm2[j] = __ldg(P2.rm + 4 * col + j); }
        }
    }
    const int items = G.N * G.chunks;
    if (active) {
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int n = item / G.chunks, ck = item % G.chunks;
            const int r_begin = ck * G.rows_per_chunk, r_end = min(G.HW, r_begin + G.rows_per_chunk);
            const long long base = (long long)n * G.HW * G.cols + col;
            const float4* pg = reinterpret_cast<const float4*>(gy) + base;
            const float4* pg2 = gy2 ? reinterpret_cast<const float4*>(gy2) + base : nullptr;
            const float4* p1 = reinterpret_cast<const float4*>(x1) + base;
            const float4* p2 = reinterpret_cast<const float4*>(r) + base;
            const uint8_t* pm = MASK ? mask + base : nullptr;
            float4* o1 = reinterpret_cast<float4*>(gx1) + base;
            float4* o2 = reinterpret_cast<float4*>(gr) + base;
            float ce[4] = {0.f, 0.f, 0.f, 0.f};
            if (ENERGY) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ce[j] = __fmul_rn(__ldg(ge + (long long)n * G.C + 4 * col + j), two_inv_hw);
            }
            for (int row = r_begin + rsub; row < r_end; row += kTailBwdDepth * G.lanes_r) {
                float4 g[kTailBwdDepth], g2[kTailBwdDepth], u[kTailBwdDepth], v[kTailBwdDepth];
                unsigned open[kTailBwdDepth];
#pragma unroll
                for (int d = 0; d < kTailBwdDepth; ++d) {
                    const int rr = row + d * G.lanes_r;
                    u[d] = v[d] = make_float4(0.f, 0.f, 0.f, 0.f);
                    open[d] = 0;
                    if (rr < r_end) {
                        g[d] = ld_stream(pg + (long long)rr * G.cols);
                        if (pg2) g2[d] = ld_stream(pg2 + (long long)rr * G.cols);
                        if (NEED_X1) u[d] = ld_stream(p1 + (long long)rr * G.cols);
                        if (NEED_R) v[d] = ld_stream(p2 + (long long)rr * G.cols);
                        if (MASK) open[d] = __ldg(pm + (long long)rr * G.cols);
                    }
                }
#pragma unroll
                for (int d = 0; d < kTailBwdDepth; ++d) {
                    const int rr = row + d * G.lanes_r;
                    if (rr < r_end) {
                        float gs[4] = {g[d].x, g[d].y, g[d].z, g[d].w};
                        if (pg2) {        // the unit's output fed two consumers: their gradients are summed here
                            gs[0] = __fadd_rn(gs[0], g2[d].x); gs[1] = __fadd_rn(gs[1], g2[d].y);
                            gs[2] = __fadd_rn(gs[2], g2[d].z); gs[3] = __fadd_rn(gs[3], g2[d].w);
                        }
                        const float xs[4] = {u[d].x, u[d].y, u[d].z, u[d].w};
                        const float rs[4] = {v[d].x, v[d].y, v[d].z, v[d].w};
                        float d1[4], d2[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float z1 = fmaf(xs[j], a1[j], b1[j]);
                            float gm;
                            if (MASK) {
                                gm = ((open[d] >> j) & 1u) ? gs[j] : 0.0f;
                            } else {
                                const float id = IDBN ? fmaf(rs[j], a2[j], b2[j]) : rs[j];
                                const float s = __fadd_rn(z1, id);
                                gm = (s <= 0.0f) ? 0.0f : gs[j];                  // aten::threshold_backward
                            }
                            const float dz = ENERGY ? __fadd_rn(gm, __fmul_rn(ce[j], z1)) : gm;
                            if (REDUCE) {
                                sb1[j] += dz; sw1[j] = fmaf(dz, xs[j] - rm1[j], sw1[j]);
                                if (IDBN) { sb2[j] += gm; sw2[j] = fmaf(gm, rs[j] - rm2[j], sw2[j]); }
                            }
                            d1[j] = dz * a1[j];
                            d2[j] = IDBN ? gm * a2[j] : gm;
                        }
                        st_out(o1 + (long long)rr * G.cols, make_float4(d1[0], d1[1], d1[2], d1[3]));
                        st_out(o2 + (long long)rr * G.cols, make_float4(d2[0], d2[1], d2[2], d2[3]));
                    }
                }
            }
        }
    }
    if (REDUCE) {
        constexpr int kSets = IDBN ? 2 : 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[threadIdx.x * 4 + j] = active ? sb1[j] : 0.f;
            red[kBThreads * 4 + threadIdx.x * 4 + j] = active ? sw1[j] : 0.f;
            if (IDBN) {
                red[2 * kBThreads * 4 + threadIdx.x * 4 + j] = active ? sb2[j] : 0.f;
                red[3 * kBThreads * 4 + threadIdx.x * 4 + j] = active ? sw2[j] : 0.f;
            }
        }
        lane_tree_fold<2 * kSets>(red, rsub, G.cols, G.lanes_r);      // fixed tree over the row-lanes
        const int Ct = kSets * G.C;              // partial[cta][Ct][2]: BN1 channels first, then BN2
        if (active && rsub == 0) {
#pragma unroll
            for (int set = 0; set < kSets; ++set) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float tb = red[(2 * set) * kBThreads * 4 + threadIdx.x * 4 + j];
                    const float tw = red[(2 * set + 1) * kBThreads * 4 + threadIdx.x * 4 + j];
                    const TailBn& P = set ? P2 : P1;
                    const float inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(P.rv + 4 * col + j), P.eps)));
                    double* p = part + ((size_t)blockIdx.x * Ct + (size_t)set * G.C + 4 * col + j) * 2;
                    p[0] = (double)tw * (double)inv;     // dW
                    p[1] = (double)tb;                   // dB
                }
            }
        }
    }
}

// Work items per image.  The grid is min(items, slots) CTAs walking the items, so the kernel takes
// ceil(items / slots) rounds: pick the split whose last round is fullest (256 images on 444 slots: one chunk
// per image fills 58 % of one round, five chunks fill 2.88 of 3), preferring fewer chunks on near-ties -- but
// only while a chunk keeps >= 8 passes of the row loop: below that the per-item epilogue (energy fold, barriers)
// costs more than the fuller round gives (measured: 14x14 and 7x7 planes lose 10 % when split).
static int make_tail_geom(int N, int C, long long HW, long long slots, int depth, TailGeom& G) {
    if (C % 4 != 0 || C / 4 > kBThreads || HW > 0x7fffffffLL || (long long)N * 16 > 0x7fffffffLL) return OODFQ_EINVAL;
    G.N = N; G.C = C; G.HW = (int)HW; G.cols = C / 4;
    G.lanes_r = kBThreads / G.cols;
    const int passes = (int)((HW + (long long)G.lanes_r * depth - 1) / ((long long)G.lanes_r * depth));
    const int cmax = passes / 8 < 1 ? 1 : (passes / 8 < 16 ? passes / 8 : 16);
    int best = 1;
    double best_eff = 0.0;
    for (int c = 1; c <= cmax; ++c) {
        const double rounds = (double)N * c / (double)slots;
        const double eff = rounds / (double)(long long)(rounds + 0.999999);
        if (eff > best_eff + 0.02) { best = c; best_eff = eff; }
    }
    G.chunks = best;
    G.rows_per_chunk = (int)((HW + best - 1) / best);
    return OODFQ_OK;
}

template <bool QUANT, bool IDBN>
static void launch_tail_fwd(bool energy, unsigned grid, cudaStream_t st, const float* x1, const float* r, float* y,
                            float* epart, uint8_t* mask, const TailGeom& G, const TailBn& P1, const TailBn& P2,
                            float inv_hw, const float* lo, const float* hi, int k) {
    if (energy) res_tail_fwd_kernel<QUANT, IDBN, true><<<grid, kBThreads, 0, st>>>(x1, r, y, epart, mask, G, P1, P2, inv_hw, lo, hi, k);
    else res_tail_fwd_kernel<QUANT, IDBN, false><<<grid, kBThreads, 0, st>>>(x1, r, y, epart, mask, G, P1, P2, inv_hw, lo, hi, k);
}

template <bool IDBN, bool ENERGY, bool REDUCE, bool MASK>
static int tail_bwd_occupancy() {
    static const int n = resident_ctas(res_tail_bwd_kernel<IDBN, ENERGY, REDUCE, MASK>, kBThreads);
    return n;
}

// (resident CTAs per SM, launcher) of one variant
struct TailBwdArgs {
    const float *gy, *gy2, *ge, *x1, *r;
    const uint8_t* mask;
    float *gx1, *gr;
    TailGeom G;
    TailBn P1, P2;
    float two_inv_hw;
    double* part;
};

template <bool IDBN, bool ENERGY, bool REDUCE, bool MASK>
static void tail_bwd_go(unsigned grid, cudaStream_t st, const TailBwdArgs& a) {
    res_tail_bwd_kernel<IDBN, ENERGY, REDUCE, MASK><<<grid, kBThreads, 0, st>>>(a.gy, a.gy2, a.ge, a.x1, a.r, a.mask, a.gx1,
                                                                              a.gr, a.G, a.P1, a.P2, a.two_inv_hw, a.part);
}

template <typename F>
static auto tail_bwd_dispatch(bool idbn, bool energy, bool reduce, bool mask, F&& f) {
#define OODFQ_TB(I, E, R, M) if (idbn == I && energy == E && reduce == R && mask == M) return f(std::integral_constant<int, (I << 3) | (E << 2) | (R << 1) | M>{});
    OODFQ_TB(0, 0, 0, 0) OODFQ_TB(0, 0, 0, 1) OODFQ_TB(0, 0, 1, 0) OODFQ_TB(0, 0, 1, 1)
    OODFQ_TB(0, 1, 0, 0) OODFQ_TB(0, 1, 0, 1) OODFQ_TB(0, 1, 1, 0) OODFQ_TB(0, 1, 1, 1)
    OODFQ_TB(1, 0, 0, 0) OODFQ_TB(1, 0, 0, 1) OODFQ_TB(1, 0, 1, 0) OODFQ_TB(1, 0, 1, 1)
    OODFQ_TB(1, 1, 0, 0) OODFQ_TB(1, 1, 0, 1) OODFQ_TB(1, 1, 1, 0)
#undef OODFQ_TB
    return f(std::integral_constant<int, 15>{});
}

}  // namespace oodfq

using namespace oodfq;

extern "C" size_t oodfq_res_tail_scratch_floats(int N, int C) { return (size_t)16 * (size_t)N * (size_t)C; }

extern "C" int oodfq_res_tail_forward(const float* x1, const float* r, float* y, float* energy, float* scratch,
                                      uint8_t* relu_mask, int N, int C, long long HW, const float* w1, const float* b1, const float* rm1,
                                      const float* rv1, float eps1, const float* w2, const float* b2,
                                      const float* rm2, const float* rv2, float eps2, int flags, const float* fq_lo,
                                      const float* fq_hi, int fq_k, oodfq_stream_t stream) {
    if (!x1 || !r || !y || !rm1 || !rv1) return fail(OODFQ_EINVAL, "res_tail_forward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "res_tail_forward: empty tensor");
    if (!(flags & OODFQ_BN_NHWC)) return fail(OODFQ_EINVAL, "res_tail_forward: channels_last only");
    const bool quant = flags & OODFQ_BN_QUANT, idbn = rv2 != nullptr;
    if (idbn && !rm2) return fail(OODFQ_EINVAL, "res_tail_forward: identity BatchNorm needs both running statistics");
    if (quant && (!fq_lo || !fq_hi || fq_k < 1 || fq_k > 8))
        return fail(OODFQ_EINVAL, "res_tail_forward: fake-quant needs a range and k in [1,8]");
    if (energy && !scratch) return fail(OODFQ_EINVAL, "res_tail_forward: the energy output needs the scratch buffer");
    static const int occ[8] = {
        resident_ctas(res_tail_fwd_kernel<false, false, false>, kBThreads), resident_ctas(res_tail_fwd_kernel<false, false, true>, kBThreads),
        resident_ctas(res_tail_fwd_kernel<false, true, false>, kBThreads), resident_ctas(res_tail_fwd_kernel<false, true, true>, kBThreads),
        resident_ctas(res_tail_fwd_kernel<true, false, false>, kBThreads), resident_ctas(res_tail_fwd_kernel<true, false, true>, kBThreads),
        resident_ctas(res_tail_fwd_kernel<true, true, false>, kBThreads), resident_ctas(res_tail_fwd_kernel<true, true, true>, kBThreads)};
    const int per_sm = occ[(quant ? 4 : 0) + (idbn ? 2 : 0) + (energy ? 1 : 0)];
    TailGeom G;
    if (make_tail_geom(N, C, HW, (long long)kNumSM * per_sm, kTailFwdDepth, G) != OODFQ_OK || !aligned16(x1) || !aligned16(r) || !aligned16(y))
        return fail(OODFQ_EINVAL, "res_tail_forward: needs C %% 4 == 0, C <= 1024 and 16-byte aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const TailBn P1{w1, b1, rm1, rv1, eps1}, P2{w2, b2, rm2, rv2, eps2};
    const long long items = (long long)N * G.chunks, cap = (long long)kNumSM * per_sm;
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    const float inv_hw = (float)(1.0 / (double)HW);
    float* epart = energy ? (G.chunks == 1 ? energy : scratch) : nullptr;
    if (quant && idbn) launch_tail_fwd<true, true>(energy, grid, st, x1, r, y, epart, relu_mask, G, P1, P2, inv_hw, fq_lo, fq_hi, fq_k);
    else if (quant) launch_tail_fwd<true, false>(energy, grid, st, x1, r, y, epart, relu_mask, G, P1, P2, inv_hw, fq_lo, fq_hi, fq_k);
    else if (idbn) launch_tail_fwd<false, true>(energy, grid, st, x1, r, y, epart, relu_mask, G, P1, P2, inv_hw, fq_lo, fq_hi, fq_k);
    else launch_tail_fwd<false, false>(energy, grid, st, x1, r, y, epart, relu_mask, G, P1, P2, inv_hw, fq_lo, fq_hi, fq_k);
    count_launch();
    int rc = check_launch("res_tail_forward");
    if (rc != OODFQ_OK || !energy || G.chunks == 1) return rc;
    const long long nc = (long long)N * C;
    res_tail_fold_energy_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(scratch, energy, nc, G.chunks);
    count_launch();
    return check_launch("res_tail_forward(fold)");
}

extern "C" int oodfq_res_tail_backward(const float* grad_y, const float* grad_y2, const float* grad_energy, const float* x1, const float* r,
                                       const uint8_t* relu_mask, float* grad_x1, float* grad_r, int N, int C, long long HW, const float* w1,
                                       const float* b1, const float* rm1, const float* rv1, float eps1,
                                       const float* w2, const float* b2, const float* rm2, const float* rv2,
                                       float eps2, int flags, float* dwdb, void* workspace, oodfq_stream_t stream) {
    if (!grad_y || !grad_x1 || !grad_r || !rm1 || !rv1) return fail(OODFQ_EINVAL, "res_tail_backward: null pointer");
    if (N <= 0 || C <= 0 || HW <= 0) return fail(OODFQ_EINVAL, "res_tail_backward: empty tensor");
    if (!(flags & OODFQ_BN_NHWC)) return fail(OODFQ_EINVAL, "res_tail_backward: channels_last only");
    const bool idbn = rv2 != nullptr, energy = grad_energy != nullptr, reduce = dwdb != nullptr, mask = relu_mask != nullptr;
    if (idbn && !rm2) return fail(OODFQ_EINVAL, "res_tail_backward: identity BatchNorm needs both running statistics");
    if (reduce && !workspace) return fail(OODFQ_EINVAL, "res_tail_backward: parameter gradients need the workspace");
    // what the chosen variant dereferences: x1 unless the mask makes it redundant, r likewise
    if (!x1 && (energy || reduce || !mask)) return fail(OODFQ_EINVAL, "res_tail_backward: x1 is needed (energy / parameter gradients / no mask)");
    if (!r && (!mask || (idbn && reduce))) return fail(OODFQ_EINVAL, "res_tail_backward: r is needed (no mask, or the identity BatchNorm's weight gradient)");
    const int per_sm = tail_bwd_dispatch(idbn, energy, reduce, mask, [](auto tag) {
        constexpr int v = decltype(tag)::value;
        return tail_bwd_occupancy<(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1>();
    });
    const int Ct = (idbn ? 2 : 1) * C;
    long long cap = (long long)kNumSM * per_sm;
    const long long table = (long long)kMaxBnSplit * kMaxBnChannels / Ct;     // rows of ws->bn_partial
    if (reduce && cap > table) cap = table;
    TailGeom G;
    if (make_tail_geom(N, C, HW, cap, kTailBwdDepth, G) != OODFQ_OK || !aligned16(grad_y) || (x1 && !aligned16(x1)) ||
        (r && !aligned16(r)) || !aligned16(grad_x1) || !aligned16(grad_r) || (grad_y2 && !aligned16(grad_y2)))
        return fail(OODFQ_EINVAL, "res_tail_backward: needs C %% 4 == 0, C <= 1024 and 16-byte aligned buffers");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    const long long items = (long long)N * G.chunks;
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    double* part = reduce ? fold_target(ws->bn_partial, Ct, (int)grid) : nullptr;     // workspace, or its own region (fold.cu)
    const TailBwdArgs args{grad_y, grad_y2, grad_energy, x1, r, relu_mask, grad_x1, grad_r, G,
                           TailBn{w1, b1, rm1, rv1, eps1}, TailBn{w2, b2, rm2, rv2, eps2}, (float)(2.0 / (double)HW), part};
    tail_bwd_dispatch(idbn, energy, reduce, mask, [&](auto tag) {
        constexpr int v = decltype(tag)::value;
        tail_bwd_go<(v >> 3) & 1, (v >> 2) & 1, (v >> 1) & 1, v & 1>(grid, st, args);
        return 0;
    });
    count_launch();
    int rc = check_launch("res_tail_backward");
    if (rc != OODFQ_OK || !reduce) return rc;
    // dwdb[0 .. Ct) = dW (BN1 channels, then BN2), dwdb[Ct .. 2 Ct) = dB
    return fold_finish(part, ws->bn_partial, Ct, (int)grid, dwdb, st);
}
