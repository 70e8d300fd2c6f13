// Calibrating QuantAct forward: data min/max -> running-range update -> fake-quant.
//
// Replaces QuantAct.forward with running_stat=True (quantization_utils/
// quant_modules.py:80-94): two full-tensor ATen reductions, ~10 scalar kernels for
// the bias-corrected EMA, then the six-pass fake-quant.  Here: one reducing kernel
// whose last CTA updates (x_min, x_max, beta_t) in place on the device, then the
// streaming fake-quant kernel walking the tensor back to front so it starts on the
// part of x the reduction left in L2.
//
// Roofline: HBM.  12 algorithmic bytes per element when 4*numel exceeds L2 (the
// quantised range depends on the min/max of the very tensor being quantised, so x
// must be read twice), 8 B/elem when x stays L2-resident between the passes.
#include <cooperative_groups.h>

#include "bn_geom.cuh"

namespace cg = cooperative_groups;

namespace oodfq {

constexpr int kRThreads = 256;
constexpr int kRUnroll = 4;

// x is read again by the quantising pass: default L2 policy here (no evict-first).
__device__ __forceinline__ float4 ld_keep(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kRThreads)
minmax_ema_kernel(const float* __restrict__ x, long long numel, Workspace* ws, float* out2,
                  float* x_min, float* x_max, const float* beta, float* beta_t, int symmetric,
                  int vec) {
    float mn = __int_as_float(0x7f800000);   // +inf
    float mx = __int_as_float(0xff800000);   // -inf
    if (vec) {
        const long long n4 = numel >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const long long tile = (long long)kRThreads * kRUnroll;
        for (long long base = (long long)blockIdx.x * tile; base < n4; base += (long long)gridDim.x * tile) {
            float4 v[kRUnroll];
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                long long i = base + u * kRThreads + threadIdx.x;
                // out-of-range slots re-read element 0: harmless for min/max
                v[u] = ld_keep(x4 + (i < n4 ? i : 0));
            }
#pragma unroll
            for (int u = 0; u < kRUnroll; ++u) {
                mn = min_nan(min_nan(mn, v[u].x), min_nan(v[u].y, min_nan(v[u].z, v[u].w)));
                mx = max_nan(max_nan(mx, v[u].x), max_nan(v[u].y, max_nan(v[u].z, v[u].w)));
            }
        }
        if (blockIdx.x == 0) {
            long long i = (n4 << 2) + threadIdx.x;
            if (i < numel) { float t = x[i]; mn = min_nan(mn, t); mx = max_nan(mx, t); }
        }
    } else {
        for (long long i = (long long)blockIdx.x * kRThreads + threadIdx.x; i < numel;
             i += (long long)gridDim.x * kRThreads) {
            float t = x[i];
            mn = min_nan(mn, t);
            mx = max_nan(mx, t);
        }
    }

    __shared__ float s_mn[kRThreads / 32], s_mx[kRThreads / 32];
    __shared__ int s_last;
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kRThreads / 32; ++w) { mn = min_nan(mn, s_mn[w]); mx = max_nan(mx, s_mx[w]); }
        ws->mm_partial[2 * blockIdx.x] = mn;
        ws->mm_partial[2 * blockIdx.x + 1] = mx;
        __threadfence();
        int t = atomicAdd(&ws->ticket[0], 1);
        s_last = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;

    // last CTA: fold the per-CTA partials, then the scalar state update
    __threadfence();
    mn = __int_as_float(0x7f800000);
    mx = __int_as_float(0xff800000);
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kRThreads) {
        mn = min_nan(mn, __ldcg(&ws->mm_partial[2 * b]));
        mx = max_nan(mx, __ldcg(&ws->mm_partial[2 * b + 1]));
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kRThreads / 32; ++w) { mn = min_nan(mn, s_mn[w]); mx = max_nan(mx, s_mx[w]); }
        ws->ticket[0] = 0;  // ready for the next launch on this stream
        if (out2) { out2[0] = mn; out2[1] = mx; }
        if (x_min) {
            if (symmetric) {   // quant_modules.py:369-374: range = +-max(|min|, |max|)
                float m = max_nan(fabsf(mn), fabsf(mx));
                mn = -m;
                mx = m;
            }
            const float b = *beta;
            const float bt = __fmul_rn(*beta_t, b);           // quant_modules.py:87
            *x_min = ema_step(*x_min, mn, b, bt);             // :88
            *x_max = ema_step(*x_max, mx, b, bt);             // :89
            *beta_t = bt;
        }
    }
}

static int launch_minmax(const float* x, long long numel, void* workspace, float* out2, float* x_min,
                         float* x_max, const float* beta, float* beta_t, int symmetric, cudaStream_t st) {
    const int vec = (aligned16(x) && numel >= 4) ? 1 : 0;
    long long per_block = vec ? (long long)kRThreads * kRUnroll * 4 : (long long)kRThreads;
    long long blocks = (numel + per_block - 1) / per_block;
    long long cap = (long long)kNumSM * 8;
    if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
    int grid = (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
    minmax_ema_kernel<<<grid, kRThreads, 0, st>>>(x, numel, reinterpret_cast<Workspace*>(workspace), out2,
                                                   x_min, x_max, beta, beta_t, symmetric, vec);
    count_launch();
    return check_launch("minmax");
}

// ---- single-pass calibrating QuantAct for tensors the chip can hold ---------------------------------------
// north_star (b): ONE pass that updates the running range, emits the fake-quantised tensor and accumulates the
// per-channel sum / sum of squares, staged through TMA and shared memory.
//
// The quantised range depends on the min/max of the very tensor being quantised, so the two-kernel path reads x
// twice.  148 SMs x 224 KB of shared memory hold 33 MB, though: a cooperative grid of one CTA per SM keeps its
// slice of x on chip while the range is reduced across the grid, and quantises from shared memory after one
// grid-wide barrier -- x crosses HBM exactly once (4 B read + 4 B written per element).  Every activation of
// the 32x32 / 28x28 configurations and the 7x7 stage of the ImageNet one fit entirely; up to ~3x the on-chip
// capacity the remainder of a slice is re-read through the 126 MB L2.  Larger tensors take the two-kernel
// path (their second read has to come from HBM whatever the kernel does).
//
// Staging: one thread arms one mbarrier per 16 KB chunk and issues all `cp.async.bulk` copies of the CTA's
// on-chip part up front; the copy engine streams x into shared memory without passing through registers while
// the threads (a) reduce the part of the slice that does not fit on chip straight from global memory and
// (b) follow the chunks in arrival order, taking min / max from shared memory.  (Round 1 kept a register-staged
// twin of this kernel; measured on a B200 the TMA-fed one is equal or faster at every shape --
// profiles/r2_microbench_calib_tma.txt -- so it is the only one left.)
//
// Arithmetic is that of minmax_ema_kernel + fq_flat_kernel: NaN-propagating min/max (order-free), the EMA of
// quant_modules.py:87-89 evaluated identically by every CTA from the state read BEFORE the barrier (CTA 0 writes
// it back after), dequantisation by the exact 2^k-entry table.
//
// Per-channel statistics (STATS != 0; what the BN-statistics hook computes with two more ATen reductions,
// trainer_direct.py:388-393).  A CTA's slice is a contiguous run of the flat storage:
//   kStatsNHWC  channels_last, C a power of two in [4, 1024]: the vector a thread meets in every chunk has the
//               same four channels (4096 % C == 0), so eight fp32 accumulators per thread, taken around a pivot
//               that is common to all threads of a channel (the first vector of that channel group in the slice,
//               read from the tile), suffice; the threads of a channel group are folded in fp64 through a 32 KB
//               staging area behind the tile, in thread order.
//   kStatsNCHW  the slice is a run of planes; after its chunks have landed the planes of one channel are summed
//               from shared memory by ONE owner, around the first element it meets: a thread per channel for planes
//               of up to 256 elements (lanes are H*W words apart: conflict-free for odd H*W, and for even H*W lane l
//               starts l elements into its plane), a warp per channel for larger planes (lanes stride the plane).
//               No cross-thread combination is needed beyond the warp's shuffle tree.
// Either way a CTA publishes fp64 (S1, S2) per channel BEFORE the one grid barrier the kernel has anyway, and
// after it the grid's warps fold the 148 partials of one value each in a fixed tree: deterministic, no atomics.
constexpr int kCThreads = 1024;
constexpr int kCTileFloats = 56 * 1024;                    // 224 KB of the 227 KB a CTA may use
constexpr long long kCoopMaxBytes = 96ll << 20;            // beyond this the two-kernel path is used
constexpr int kTChunkVec = 1024;                           // 128-bit vectors per bulk copy (16 KB)
constexpr int kTChunksMax = kCTileFloats / (4 * kTChunkVec);   // 14
constexpr int kStatsNone = 0, kStatsNHWC = 1, kStatsNCHW = 2;
constexpr int kStageFloats = kCThreads * 8;                // kStatsNHWC: 8 fp32 partial sums per thread (32 KB)
constexpr int kStatsMaxC = 1024;
constexpr int kThreadPlaneMax = 256;                       // kStatsNCHW: planes up to this size are walked by one thread
static_assert(kTChunkVec == kCThreads, "one vector per thread and chunk");

__device__ __forceinline__ void minmax4(float& mn, float& mx, const float4& v) {
    mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
    mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
}
__device__ __forceinline__ void accum4(float (&a1)[4], float (&a2)[4], const float4& v, const float4& pv) {
    float d;
    d = v.x - pv.x; a1[0] += d; a2[0] = fmaf(d, d, a2[0]);
    d = v.y - pv.y; a1[1] += d; a2[1] = fmaf(d, d, a2[1]);
    d = v.z - pv.z; a1[2] += d; a2[2] = fmaf(d, d, a2[2]);
    d = v.w - pv.w; a1[3] += d; a2[3] = fmaf(d, d, a2[3]);
}

template <int STATS>
__global__ void __launch_bounds__(kCThreads, 1)
act_calib_onchip_kernel(const float* __restrict__ x, float* __restrict__ y, long long numel, Workspace* ws,
                        float* x_min, float* x_max, const float* beta, float* beta_t, int k, int cap_chunks,
                        int C, int HW, double* __restrict__ sums) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bars[kTChunksMax];
    __shared__ float lut[kLutMax];
    __shared__ float s_mn[kCThreads / 32], s_mx[kCThreads / 32];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // contiguous slice of this CTA, a multiple of 4 elements (x and y are 16-byte aligned, numel % 4 == 0)
    const long long n4 = numel >> 2;
    const long long per = (n4 + gridDim.x - 1) / gridDim.x;
    const long long b4 = (long long)blockIdx.x * per;
    const long long e4 = b4 + per < n4 ? b4 + per : n4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    float4* t4 = reinterpret_cast<float4*>(tile);
    const long long cap4 = (long long)cap_chunks * kTChunkVec;
    const long long len4 = e4 > b4 ? e4 - b4 : 0;
    const int on4 = (int)(len4 < cap4 ? len4 : cap4);                // vectors of the slice held on chip
    const int nch = (on4 + kTChunkVec - 1) / kTChunkVec;
    // the state every CTA needs after the barrier, read before CTA 0 may overwrite it
    const float old_min = *x_min, old_max = *x_max, b = *beta, old_bt = *beta_t;

    if (threadIdx.x == 0) {
        for (int c = 0; c < nch; ++c) mbar_init(&bars[c], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < nch; ++c) {
            const int vecs = on4 - c * kTChunkVec < kTChunkVec ? on4 - c * kTChunkVec : kTChunkVec;
            const uint32_t bytes = (uint32_t)vecs * 16u;
            mbar_arrive_expect_tx(&bars[c], bytes);
            bulk_g2s(t4 + c * kTChunkVec, x4 + b4 + c * kTChunkVec, bytes, &bars[c]);
        }
    }

    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
    const int G = C >> 2;                                            // kStatsNHWC: vectors per pixel row
    if (STATS == kStatsNHWC && nch > 0) {
        // pivot of this thread's four channels: the first vector of its channel group in the slice (chunk 0)
        mbar_wait(&bars[0], 0);
        if ((int)threadIdx.x < on4) pv = t4[threadIdx.x & (G - 1)];
    }
    constexpr int kU = 4;
    // the part that does not fit on chip (tensors above 33 MB): through registers while the copies are in flight
    for (long long base = b4 + on4; base < e4; base += (long long)kU * kCThreads) {
        float4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long i = base + u * kCThreads + threadIdx.x;
            if (i < e4) v[u] = ld_keep(x4 + i);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long i = base + u * kCThreads + threadIdx.x;
            if (i < e4) {
                minmax4(mn, mx, v[u]);
                if (STATS == kStatsNHWC) accum4(a1, a2, v[u], pv);
            }
        }
    }
    // the on-chip part, chunk by chunk as the copies land (one vector per thread and chunk)
    for (int c = 0; c < nch; ++c) {
        mbar_wait(&bars[c], 0);
        const int i = c * kTChunkVec + threadIdx.x;
        if (i < on4) {
            const float4 v = t4[i];
            minmax4(mn, mx, v);
            if (STATS == kStatsNHWC) accum4(a1, a2, v, pv);
        }
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }

    if (STATS == kStatsNHWC) {
        // fold the threads of one channel group (thread t holds group (b4 + t) mod G), in thread order, in fp64
        float* stage = tile + (size_t)cap_chunks * kTChunkVec * 4;
        float4* st4 = reinterpret_cast<float4*>(stage);
        st4[2 * threadIdx.x] = make_float4(a1[0], a1[1], a1[2], a1[3]);
        st4[2 * threadIdx.x + 1] = make_float4(a2[0], a2[1], a2[2], a2[3]);
        __syncthreads();
        const int bg = (int)(b4 & (G - 1));
        for (int ch = threadIdx.x; ch < C; ch += kCThreads) {
            const int g = ch >> 2, j = ch & 3;
            const int t0 = (g - bg + G) & (G - 1);                   // first slice position of that group
            double s1 = 0.0, s2 = 0.0;
            if (t0 < len4) {
                double t1 = 0.0, t2 = 0.0;
                for (int t = t0; t < kCThreads; t += G) {
                    t1 += (double)stage[8 * t + j];
                    t2 += (double)stage[8 * t + 4 + j];
                }
                const double cnt = (double)((len4 - 1 - t0) / G + 1);
                const double p = (double)tile[4 * t0 + j];
                s1 = t1 + cnt * p;
                s2 = t2 + 2.0 * p * t1 + cnt * p * p;
            }
            double* pp = ws->bn_partial + ((size_t)blockIdx.x * C + ch) * 2;
            pp[0] = s1;
            pp[1] = s2;
        }
    }
    if (STATS == kStatsNCHW) {
        // every chunk has landed (each thread waited on all of them)
        __syncthreads();
        // The slice is a run of planes (the first and the last possibly cut); relative plane j belongs to channel
        // (c_first + j) mod C.  All offsets below are relative to the slice start and fit 32 bits.
        const int len = (int)(len4 << 2), on = on4 << 2;
        const long long B = b4 << 2;
        const long long p_first = B / HW;
        const int rel0 = (int)(p_first * HW - B);                    // start of plane 0 relative to the slice, <= 0
        const int nplanes = len > 0 ? (len - rel0 + HW - 1) / HW : 0;
        const int c_first = (int)(p_first % C);
        const float* xs = x + B;
        if (HW <= kThreadPlaneMax) {
            // small planes: thread t owns the planes t, t + C, ... (one channel), walking each serially.  Lanes are
            // HW words apart: conflict-free for odd HW; for even HW lane l starts l elements into its plane.
            const int rot = (HW & 1) ? 0 : lane;
            for (int j0 = threadIdx.x; j0 < C; j0 += kCThreads) {
                float t1 = 0.f, t2 = 0.f, u1 = 0.f, u2 = 0.f, pivot = 0.f;
                int cnt = 0;
                for (int j = j0; j < nplanes; j += C) {
                    const int ps = rel0 + j * HW;
                    const int lo = ps > 0 ? ps : 0, hi = ps + HW < len ? ps + HW : len;
                    const int m = hi - lo;
                    if (j == j0) pivot = lo < on ? tile[lo] : __ldg(xs + lo);
                    int e = rot < m ? rot : rot % m;
                    int s = 0;
                    for (; s + 1 < m; s += 2) {
                        const int i0 = lo + e;
                        e = e + 1 < m ? e + 1 : 0;
                        const int i1 = lo + e;
                        e = e + 1 < m ? e + 1 : 0;
                        const float d0 = (i0 < on ? tile[i0] : __ldg(xs + i0)) - pivot;
                        const float d1 = (i1 < on ? tile[i1] : __ldg(xs + i1)) - pivot;
                        t1 += d0; t2 = fmaf(d0, d0, t2);
                        u1 += d1; u2 = fmaf(d1, d1, u2);
                    }
                    if (s < m) {
                        const int i0 = lo + e;
                        const float d0 = (i0 < on ? tile[i0] : __ldg(xs + i0)) - pivot;
                        t1 += d0; t2 = fmaf(d0, d0, t2);
                    }
                    cnt += m;
                }
                t1 += u1;
                t2 += u2;
                const int c = c_first + j0 < C ? c_first + j0 : c_first + j0 - C;
                const double pd = (double)pivot, cd = (double)cnt;
                double* pp = ws->bn_partial + ((size_t)blockIdx.x * C + c) * 2;
                pp[0] = (double)t1 + cd * pd;
                pp[1] = (double)t2 + 2.0 * pd * (double)t1 + cd * pd * pd;
            }
        } else {
            // large planes: warp w owns the channels of relative planes w, w + 32, ... < C; lanes stride a plane
            for (int j0 = warp; j0 < C; j0 += kCThreads / 32) {
                float t1 = 0.f, t2 = 0.f, u1 = 0.f, u2 = 0.f, pivot = 0.f;
                int cnt = 0;
                for (int j = j0; j < nplanes; j += C) {
                    const int ps = rel0 + j * HW;
                    const int lo = ps > 0 ? ps : 0, hi = ps + HW < len ? ps + HW : len;
                    if (j == j0) pivot = lo < on ? tile[lo] : __ldg(xs + lo);
                    int i = lo + lane;
                    for (; i + 32 < hi; i += 64) {
                        const float d0 = (i < on ? tile[i] : ld_stream(xs + i)) - pivot;
                        const float d1 = (i + 32 < on ? tile[i + 32] : ld_stream(xs + i + 32)) - pivot;
                        t1 += d0; t2 = fmaf(d0, d0, t2);
                        u1 += d1; u2 = fmaf(d1, d1, u2);
                    }
                    if (i < hi) {
                        const float d0 = (i < on ? tile[i] : ld_stream(xs + i)) - pivot;
                        t1 += d0; t2 = fmaf(d0, d0, t2);
                    }
                    cnt += hi - lo;
                }
                t1 = warp_sum(t1 + u1);
                t2 = warp_sum(t2 + u2);
                if (lane == 0) {
                    const int c = c_first + j0 < C ? c_first + j0 : c_first + j0 - C;
                    const double pd = (double)pivot, cd = (double)cnt;
                    double* pp = ws->bn_partial + ((size_t)blockIdx.x * C + c) * 2;
                    pp[0] = (double)t1 + cd * pd;
                    pp[1] = (double)t2 + 2.0 * pd * (double)t1 + cd * pd * pd;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kCThreads / 32; ++w) { mn = min_nan(mn, s_mn[w]); mx = max_nan(mx, s_mx[w]); }
        ws->mm_partial[2 * blockIdx.x] = mn;
        ws->mm_partial[2 * blockIdx.x + 1] = mx;
    }
    grid.sync();

    // every CTA folds the grid's partials and takes the same EMA step
    mn = __int_as_float(0x7f800000);
    mx = __int_as_float(0xff800000);
    for (int c = threadIdx.x; c < (int)gridDim.x; c += kCThreads) {
        mn = min_nan(mn, __ldcg(&ws->mm_partial[2 * c]));
        mx = max_nan(mx, __ldcg(&ws->mm_partial[2 * c + 1]));
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    __syncthreads();
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    mn = s_mn[0];
    mx = s_mx[0];
#pragma unroll
    for (int w = 1; w < kCThreads / 32; ++w) { mn = min_nan(mn, s_mn[w]); mx = max_nan(mx, s_mx[w]); }
    const float bt = __fmul_rn(old_bt, b);                       // quant_modules.py:87
    const float new_min = ema_step(old_min, mn, b, bt);          // :88
    const float new_max = ema_step(old_max, mx, b, bt);          // :89
    if (blockIdx.x == 0 && threadIdx.x == 0) { *x_min = new_min; *x_max = new_max; *beta_t = bt; }

    const QParams qp = make_qparams(new_min, new_max, k);
    const int qh = 1 << (k - 1), qmask = (1 << k) - 1;
    build_lut(lut, qp, k, threadIdx.x, kCThreads);
    __syncthreads();
    for (long long base = b4; base < e4; base += (long long)kU * kCThreads) {
        float4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long i = base + u * kCThreads + threadIdx.x;
            if (i < e4) v[u] = (i - b4 < on4) ? t4[i - b4] : ld_stream(x4 + i);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const long long i = base + u * kCThreads + threadIdx.x;
            if (i < e4) {
                float4 o;
                o.x = fake_quant_lut(v[u].x, qp, lut, qh, qmask);
                o.y = fake_quant_lut(v[u].y, qp, lut, qh, qmask);
                o.z = fake_quant_lut(v[u].z, qp, lut, qh, qmask);
                o.w = fake_quant_lut(v[u].w, qp, lut, qh, qmask);
                st_out(y4 + i, o);
            }
        }
    }
    if (STATS != kStatsNone) {
        // the grid's warps fold one value each: 148 partials, lanes over CTAs, fixed shuffle tree
        const int gw = blockIdx.x * (kCThreads / 32) + warp;
        for (int c = gw; c < C; c += gridDim.x * (kCThreads / 32))
            fold_partials(ws->bn_partial, C, c, (int)gridDim.x, lane, sums);
    }
}

struct CoopProbe {                   // per kernel variant: can it be launched cooperatively here, and how wide
    int state = 0;                   // 0 = not probed, 1 = usable, -1 = unavailable on this device / driver
    int grid = 0;
};

static const void* onchip_kernel(int stats) {
    return stats == kStatsNHWC ? (const void*)act_calib_onchip_kernel<kStatsNHWC>
         : stats == kStatsNCHW ? (const void*)act_calib_onchip_kernel<kStatsNCHW>
                               : (const void*)act_calib_onchip_kernel<kStatsNone>;
}

// true if the launch was made (the caller falls back to the two-kernel path otherwise)
static bool try_onchip_calib(const float* x, float* y, long long numel, void* workspace, float* x_min, float* x_max,
                             const float* beta, float* beta_t, int k, int stats, int C, int HW, double* sums,
                             cudaStream_t st, int* rc) {
    static CoopProbe probes[3];
    CoopProbe& pr = probes[stats];
    const void* kernel = onchip_kernel(stats);
    if (pr.state < 0 || k > 8 || (numel & 3) || numel * 4 > kCoopMaxBytes || !aligned16(x) || !aligned16(y)) return false;
    if (stats == kStatsNHWC && (C < 4 || C > kStatsMaxC || (C & (C - 1)))) return false;
    if (stats == kStatsNCHW && (C < 1 || C > kStatsMaxC || HW < 1)) return false;
    // small planes are walked one thread per plane: fine from shared memory, hopeless on the part of a slice that is
    // re-read from global memory (one sector per lane), so those tensors must fit on chip entirely
    if (stats == kStatsNCHW && HW <= kThreadPlaneMax && numel > (long long)kNumSM * kCTileFloats) return false;
    const size_t smem = (size_t)kCTileFloats * sizeof(float);
    // the statistics staging area of the channels_last variant is carved out of the same 224 KB
    int cap_chunks = stats == kStatsNHWC ? (kCTileFloats - kStageFloats) / (4 * kTChunkVec) : kTChunksMax;
    if (pr.state == 0) {
        int dev = 0, coop = 0, per_sm = 0, sms = 0;
        bool ok = cudaGetDevice(&dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
                  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kCThreads, smem) == cudaSuccess &&
                  per_sm >= 1;
        (void)cudaGetLastError();
        pr.grid = ok ? (sms < kMaxBnSplit * 4 ? sms : kMaxBnSplit * 4) : 0;
        pr.state = ok ? 1 : -1;
        if (!ok) return false;
    }
    // fp64 partials [grid][C][2] live in the workspace's bn_partial area
    if (stats != kStatsNone && (size_t)pr.grid * (size_t)C > (size_t)kMaxBnSplit * kMaxBnChannels) return false;
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    void* args[] = {(void*)&x, (void*)&y, (void*)&numel, (void*)&ws, (void*)&x_min, (void*)&x_max, (void*)&beta,
                    (void*)&beta_t, (void*)&k, (void*)&cap_chunks, (void*)&C, (void*)&HW, (void*)&sums};
    const cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(pr.grid), dim3(kCThreads), args, smem, st);
    if (e != cudaSuccess) {          // e.g. the grid cannot be co-resident right now: leave it to the other path
        (void)cudaGetLastError();
        return false;
    }
    count_launch();
    *rc = check_launch(stats ? "act_calib_stats_forward(on-chip)" : "act_calib_forward(on-chip)");
    return true;
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_minmax(const float* x, long long numel, float* out2, void* workspace,
                            oodfq_stream_t stream) {
    if (!x || !out2 || !workspace) return fail(OODFQ_EINVAL, "minmax: null pointer");
    if (numel <= 0) return fail(OODFQ_EINVAL, "minmax: empty tensor has no min/max");
    return launch_minmax(x, numel, workspace, out2, nullptr, nullptr, nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int oodfq_act_calib_forward(const float* x, float* y, int8_t* codes, long long numel,
                                       float* x_min, float* x_max, const float* beta, float* beta_t,
                                       int k, int flags, void* workspace, oodfq_stream_t stream) {
    if (!x || !x_min || !x_max || !beta || !beta_t || !workspace)
        return fail(OODFQ_EINVAL, "act_calib_forward: null pointer");
    if (numel <= 0) return fail(OODFQ_EINVAL, "act_calib_forward: empty tensor has no min/max");
    if (k < 1 || k > 16) return fail(OODFQ_EINVAL, "act_calib_forward: k=%d outside [1,16]", k);
    if (codes && k > 8) return fail(OODFQ_EINVAL, "act_calib_forward: int8 codes need k <= 8");
    const bool sym = (flags & OODFQ_SYMMETRIC) != 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = OODFQ_OK;
    if (y && !codes && !sym && !(flags & OODFQ_NO_ONCHIP) &&
        try_onchip_calib(x, y, numel, workspace, x_min, x_max, beta, beta_t, k, kStatsNone, 0, 0, nullptr, st, &rc))
        return rc;
    rc = launch_minmax(x, numel, workspace, nullptr, x_min, x_max, beta, beta_t, sym ? 1 : 0, st);
    if (rc != OODFQ_OK || !y) return rc;
    return launch_fakequant_scalar(x, y, codes, numel, x_min, x_max, k, sym, /*reverse=*/true, st);
}

extern "C" int oodfq_act_calib_stats_forward(const float* x, float* y, int N, int C, long long HW,
                                             float* x_min, float* x_max, const float* beta, float* beta_t,
                                             int k, int flags, double* sums, void* workspace,
                                             oodfq_stream_t stream) {
    if (!x || !y || !x_min || !x_max || !beta || !beta_t || !sums || !workspace)
        return fail(OODFQ_EINVAL, "act_calib_stats_forward: null pointer");
    if (N < 1 || C < 1 || HW < 1) return fail(OODFQ_EINVAL, "act_calib_stats_forward: empty tensor");
    if (k < 1 || k > 8) return fail(OODFQ_EINVAL, "act_calib_stats_forward: k=%d outside [1,8]", k);
    if (flags & OODFQ_SYMMETRIC) return fail(OODFQ_EINVAL, "act_calib_stats_forward: asymmetric ranges only");
    const bool nhwc = (flags & OODFQ_BN_NHWC) != 0;
    if (nhwc && (C & 3)) return fail(OODFQ_EINVAL, "act_calib_stats_forward: channels_last needs C %% 4 == 0");
    const long long numel = (long long)N * C * HW;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = OODFQ_OK;
    if (!(flags & OODFQ_NO_ONCHIP) && HW <= 0x7fffffff &&
        try_onchip_calib(x, y, numel, workspace, x_min, x_max, beta, beta_t, k, nhwc ? kStatsNHWC : kStatsNCHW, C,
                         (int)HW, sums, st, &rc))
        return rc;
    // tensors the chip cannot hold (or a geometry the single-pass kernel does not take): the range needs one read
    // of its own, the second read quantises AND accumulates the channel sums -- 12 B/elem, the floor above L2 size
    rc = launch_minmax(x, numel, workspace, nullptr, x_min, x_max, beta, beta_t, 0, st);
    if (rc != OODFQ_OK) return rc;
    return oodfq_bn_stats_forward(x, N, C, HW, nullptr, sums, y, x_min, x_max, k, nhwc ? OODFQ_BN_NHWC : 0, workspace,
                                  stream);
}
