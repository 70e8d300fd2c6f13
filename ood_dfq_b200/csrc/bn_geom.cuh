// Shared decomposition of an NCHW tensor for the per-channel kernels (bn_stats.cu, bn_fused.cu).
//
//   plane kernels  (H*W >= 1024, 16-byte aligned): a CTA owns one channel and a subset of the
//              batch index and streams its planes.
//   group kernels  (small planes: 7x7, 14x14 ...; or unaligned): the tensor is [N][C*HW]; a CTA
//              owns a contiguous SPAN of the inner axis covering several whole channels (or a
//              chunk of one) and a subset of n.  A thread keeps the same offsets inside the span
//              for every n, so the channel of each of its elements is fixed and per-channel state
//              stays in registers; fully coalesced even for 49-element planes.  kDepth rows of the
//              batch are in flight per thread.
// Grids are `base * split` CTAs of equal work sized to fill the resident CTA slots exactly once.
#pragma once

#include <cstdio>

#include "common.cuh"

namespace oodfq {

constexpr int kBThreads = 256;
constexpr int kSpanMax = kBThreads * 4;             // group kernels: 1024 elements = 4 KB per n and CTA
constexpr long long kPlaneMin = 1024;               // planes at least this large take the plane kernels

constexpr int kDepth = 4;                           // batch rows in flight per thread (group kernels)

// ---- eval-mode BatchNorm as a per-channel affine (shared by bn_fused.cu and the statistics kernels) -----------
struct BnParams {
    const float* w;    // [C] or NULL (1)
    const float* b;    // [C] or NULL (0)
    const float* rm;   // [C] running mean
    const float* rv;   // [C] running var
    float eps;
};

// one rounding per step, the same in every kernel (forward and the mask recomputed by backward)
__device__ __forceinline__ void affine_of(const BnParams& P, int c, float& a, float& b, float& invstd) {
    invstd = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(P.rv + c), P.eps)));
    a = __fmul_rn(P.w ? __ldg(P.w + c) : 1.0f, invstd);
    b = __fsub_rn(P.b ? __ldg(P.b + c) : 0.0f, __fmul_rn(__ldg(P.rm + c), a));
}

// clamp_min(z, 0) keeps NaN
__device__ __forceinline__ float relu_keep_nan(float z) { return (z != z) ? z : fmaxf(z, 0.0f); }

// `zr` receives the PRE-activation value a*x + b (what the ReLU mask and the debug output derive from);
// `lowc` = relu_lower_bound(qp), only read for RELU && QUANT
template <bool RELU, bool QUANT>
__device__ __forceinline__ float head(float x, float a, float b, const QParams& qp, float lowc, const float* lut, int qh,
                                      int qmask, float& zr) {
    zr = fmaf(x, a, b);
    if (RELU && QUANT) return relu_fake_quant_lut(zr, qp, lowc, lut, qh, qmask);
    const float z = RELU ? relu_keep_nan(zr) : zr;
    return QUANT ? fake_quant_lut(z, qp, lut, qh, qmask) : z;
}

// the last CTA of a channel folds that channel's `nparts` partials (one warp, fixed tree)
// (OutT = double for the BN-input statistics, which cross NVLink as fp64; float for parameter gradients, which torch
// consumes as fp32 -- rounding once here spares a conversion launch per backward)
template <typename OutT>
__device__ __forceinline__ void fold_partials(const double* partial, int C, int c, int nparts, int lane,
                                              OutT* sums) {
    double t1 = 0.0, t2 = 0.0;
    for (int p = lane; p < nparts; p += 32) {
        const double2 q = __ldcg(reinterpret_cast<const double2*>(partial + ((size_t)p * C + c) * 2));
        t1 += q.x;
        t2 += q.y;
    }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    if (lane == 0) {
        sums[c] = (OutT)t1;
        sums[C + c] = (OutT)t2;
    }
}

// ---- channels_last reducing kernels: CTA-level fold of the row-lanes ------------------------------------------
// Every thread of a channels_last kernel holds partial sums for its 4 channels; thread t = rsub * wcols + lc shares
// its column lc with the threads rsub' * wcols + lc.  lane_tree_fold adds them in a fixed binary tree over rsub
// (deterministic; log2(lanes_r) barriers) and leaves the totals in the rsub == 0 entries of `red` -- the serial
// loop it replaces cost 4-5 us for 16-channel tensors (4 columns, 64 row-lanes, 4 threads doing all the adding).
//   red: [NV][kBThreads][4] floats; every thread has stored its NV x 4 values at red[(v * kBThreads + t) * 4 + j]
template <int NV>
__device__ __forceinline__ void lane_tree_fold(float* red, int rsub, int wcols, int lanes_r) {
    int s = 1;
    while (s * 2 < lanes_r) s *= 2;                  // largest power of two below lanes_r (lanes_r >= 2)
    for (; s > 0; s >>= 1) {
        __syncthreads();
        if (rsub < s && rsub + s < lanes_r) {
            const int t = threadIdx.x, u = t + s * wcols;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float4* mine = reinterpret_cast<float4*>(red + ((size_t)v * kBThreads + t) * 4);
                const float4 o = *reinterpret_cast<const float4*>(red + ((size_t)v * kBThreads + u) * 4);
                float4 m = *mine;
                m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
                *mine = m;
            }
        }
    }
    __syncthreads();
}

struct BnGeom {
    int N, C;
    long long HW;
    int cg;          // channels per group (1 when a plane is split into chunks)
    int groups;      // ceil(C / cg)
    int chunks;      // chunks per plane (1 when cg > 1)
    long long chunk_len;   // elements per chunk (multiple of VEC)
    int split;       // CTAs along n
};

// span of CTA (g, ck): element offset inside the [C*HW] row, its length and first channel
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
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
    if (VEC == 4) {
        float4 t = ld_stream(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
    } else {
        v[0] = ld_stream(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
    if (VEC == 4) st_out(reinterpret_cast<float4*>(p), make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]));
    else *p = v[0];
}

// ---- channels_last ---------------------------------------------------------------------------
// x is [R = N*H*W rows][C]; channels are the fastest axis.  A thread owns one 128-bit column
// (4 consecutive channels, fixed for its lifetime -> per-channel state in registers) and walks down
// the rows; the CTA's 256 threads cover `lanes_r = 256 / (C/4)` rows at a time as ONE contiguous run.
struct NhwcGeom {
    long long R;        // rows
    int C, cols;        // channels, 128-bit columns per row (C/4)
    int lanes_r;        // rows covered by one pass of the CTA (>= 1); threads >= lanes_r*min(cols,256) idle
    int col_blocks;     // ceil(cols / 256) when a row is wider than the CTA
};

__host__ __device__ inline NhwcGeom make_nhwc(long long R, int C) {
    NhwcGeom G;
    G.R = R; G.C = C; G.cols = C / 4;
    G.lanes_r = G.cols <= kBThreads ? kBThreads / G.cols : 1;
    G.col_blocks = (G.cols + kBThreads - 1) / kBThreads;
    return G;
}

// every CTA left one fp64 partial pair per channel in partial[cta][C][2]; sum them in CTA order
template <typename OutT>
static __global__ void __launch_bounds__(kBThreads)
bn_nhwc_fold_kernel(const double* __restrict__ partial, int C, int nparts, OutT* __restrict__ dwdb) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (kBThreads / 32) + (threadIdx.x >> 5);
    if (c < C) fold_partials(partial, C, c, nparts, lane, dwdb);
}


// ---- host side ----------------------------------------------------------------------------
// fold.cu: where a reducing kernel writes its per-CTA partials (the workspace, or its own arena region while folds
// are deferred) and what happens behind it (an immediate fold launch, or a note for the deferred multi-tensor fold)
double* fold_target(double* ws_partial, int C, int nparts);
int fold_finish(double* target, double* ws_partial, int C, int nparts, float* out, cudaStream_t st);
int fold_finish(double* target, double* ws_partial, int C, int nparts, double* out, cudaStream_t st);

inline bool plane_ok(long long HW, bool vec_ok) { return vec_ok && HW >= kPlaneMin && (HW % 4) == 0; }

// CTAs along the batch axis.  The grid is `base * split` CTAs of equal work, so it should fill the
// resident slots of the machine exactly once (a grid of 1.3 waves idles most SMs for the last 0.3).
inline int pick_split(long long base, int N, long long cap, int slots) {
    long long s = slots / base;
    if (s > cap) s = cap;
    if (s > N) s = N;
    return (int)(s < 1 ? 1 : s);
}

inline int make_geom(int N, int C, long long HW, bool vec_ok, int slots, BnGeom& G, int& vec, char* why, size_t whyn) {
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
            if (cg < 2 && C >= 2) {          // cg == 1 means "chunk mode" to the kernels
                G.cg = 1; G.groups = C; G.chunks = 1; G.chunk_len = HW;
            } else {
                G.cg = (int)cg;
                G.groups = (C + G.cg - 1) / G.cg;
                G.chunks = 1;
                G.chunk_len = (long long)G.cg * HW;
            }
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
    G.split = pick_split((long long)G.groups * G.chunks, N, kMaxBnSplit / G.chunks, slots);
    return OODFQ_OK;
}

}  // namespace oodfq
