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

#include "common.cuh"

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
// The quantised range depends on the min/max of the very tensor being quantised, so the two-kernel path reads x
// twice.  148 SMs x 224 KB of shared memory hold 33 MB, though: a cooperative grid of one CTA per SM keeps its
// slice of x on chip while the range is reduced across the grid, and quantises from shared memory after one
// grid-wide barrier -- x crosses HBM exactly once (4 B read + 4 B written per element).  Every activation of
// the 32x32 / 28x28 configurations and the 7x7 stage of the ImageNet one fit entirely; up to ~3x the on-chip
// capacity the remainder of a slice is re-read through the 126 MB L2.  Larger tensors take the two-kernel
// path (their second read has to come from HBM whatever the kernel does).
//
// Arithmetic is that of minmax_ema_kernel + fq_flat_kernel: NaN-propagating min/max (order-free), the EMA of
// quant_modules.py:87-89 evaluated identically by every CTA from the state read BEFORE the barrier (CTA 0 writes
// it back after), dequantisation by the exact 2^k-entry table.
constexpr int kCThreads = 1024;
constexpr int kCTileFloats = 56 * 1024;                    // 224 KB of the 227 KB a CTA may use
constexpr long long kCoopMaxBytes = 96ll << 20;            // beyond this the two-kernel path is used

__global__ void __launch_bounds__(kCThreads, 1)
act_calib_onchip_kernel(const float* __restrict__ x, float* __restrict__ y, long long numel, Workspace* ws,
                        float* x_min, float* x_max, const float* beta, float* beta_t, int k) {
    extern __shared__ __align__(16) float tile[];
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
    const long long cap4 = kCTileFloats / 4;
    // the state every CTA needs after the barrier, read before CTA 0 may overwrite it
    const float old_min = *x_min, old_max = *x_max, b = *beta, old_bt = *beta_t;

    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    constexpr int kU = 4;                                        // 128-bit loads in flight per thread
    for (long long base = b4; base < e4; base += (long long)kU * kCThreads) {
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
                if (i - b4 < cap4) t4[i - b4] = v[u];
                mn = min_nan(min_nan(mn, v[u].x), min_nan(v[u].y, min_nan(v[u].z, v[u].w)));
                mx = max_nan(max_nan(mx, v[u].x), max_nan(v[u].y, max_nan(v[u].z, v[u].w)));
            }
        }
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
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
            if (i < e4) v[u] = (i - b4 < cap4) ? t4[i - b4] : ld_stream(x4 + i);
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
}


// ---- the same kernel with the on-chip part of the slice fetched by TMA bulk copies (north_star (b): "staged
// through TMA / shared memory") -------------------------------------------------------------------------------
// One thread arms one mbarrier per 16 KB chunk and issues all `cp.async.bulk` copies of the CTA's on-chip part up
// front; the copy engine streams x into shared memory without passing through registers while the threads
// (a) reduce the part of the slice that does not fit on chip straight from global memory and (b) follow the
// chunks in arrival order, taking min / max from shared memory.  Everything after the grid barrier is the
// register-staged kernel's.  OPT-IN (OODFQ_ONCHIP_TMA) until it has been measured against that kernel on a B200:
// it was written in a session without GPU minutes.
constexpr int kTChunkVec = 1024;                                   // 128-bit vectors per bulk copy (16 KB)
constexpr int kTChunks = kCTileFloats / (4 * kTChunkVec);          // 14
static_assert(kTChunkVec == kCThreads, "one vector per thread and chunk");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" :: "r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" :: "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy (size and both addresses multiples of 16 bytes), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a byte count that never completes must end in a trap (a CUDA error), not in a hung device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t a = smem_u32(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(phase) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}

__global__ void __launch_bounds__(kCThreads, 1)
act_calib_onchip_tma_kernel(const float* __restrict__ x, float* __restrict__ y, long long numel, Workspace* ws,
                            float* x_min, float* x_max, const float* beta, float* beta_t, int k) {
    extern __shared__ __align__(128) float tile_tma[];
    __shared__ __align__(8) uint64_t bars[kTChunks];
    __shared__ float lut[kLutMax];
    __shared__ float s_mn[kCThreads / 32], s_mx[kCThreads / 32];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n4 = numel >> 2;
    const long long per = (n4 + gridDim.x - 1) / gridDim.x;
    const long long b4 = (long long)blockIdx.x * per;
    const long long e4 = b4 + per < n4 ? b4 + per : n4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    float4* t4 = reinterpret_cast<float4*>(tile_tma);
    const long long cap4 = kCTileFloats / 4;
    const long long len4 = e4 > b4 ? e4 - b4 : 0;
    const int on4 = (int)(len4 < cap4 ? len4 : cap4);                // vectors of the slice held on chip
    const int nch = (on4 + kTChunkVec - 1) / kTChunkVec;
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
                mn = min_nan(min_nan(mn, v[u].x), min_nan(v[u].y, min_nan(v[u].z, v[u].w)));
                mx = max_nan(max_nan(mx, v[u].x), max_nan(v[u].y, max_nan(v[u].z, v[u].w)));
            }
        }
    }
    // the on-chip part, chunk by chunk as the copies land (one vector per thread and chunk)
    for (int c = 0; c < nch; ++c) {
        mbar_wait(&bars[c], 0);
        const int i = c * kTChunkVec + threadIdx.x;
        if (i < on4) {
            const float4 v = t4[i];
            mn = min_nan(min_nan(mn, v.x), min_nan(v.y, min_nan(v.z, v.w)));
            mx = max_nan(max_nan(mx, v.x), max_nan(v.y, max_nan(v.z, v.w)));
        }
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kCThreads / 32; ++w) { mn = min_nan(mn, s_mn[w]); mx = max_nan(mx, s_mx[w]); }
        ws->mm_partial[2 * blockIdx.x] = mn;
        ws->mm_partial[2 * blockIdx.x + 1] = mx;
    }
    grid.sync();

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
}

struct CoopProbe {                   // per kernel variant: can it be launched cooperatively here, and how wide
    int state = 0;                   // 0 = not probed, 1 = usable, -1 = unavailable on this device / driver
    int grid = 0;
};

// true if the launch was made (the caller falls back to the two-kernel path otherwise)
static bool try_onchip_calib(const float* x, float* y, long long numel, void* workspace, float* x_min, float* x_max,
                             const float* beta, float* beta_t, int k, bool tma, cudaStream_t st, int* rc) {
    static CoopProbe probes[2];
    CoopProbe& pr = probes[tma ? 1 : 0];
    const void* kernel = tma ? (const void*)act_calib_onchip_tma_kernel : (const void*)act_calib_onchip_kernel;
    if (pr.state < 0 || k > 8 || (numel & 3) || numel * 4 > kCoopMaxBytes || !aligned16(x) || !aligned16(y)) return false;
    const size_t smem = (size_t)kCTileFloats * sizeof(float);
    if (pr.state == 0) {
        int dev = 0, coop = 0, per_sm = 0, sms = 0;
        bool ok = cudaGetDevice(&dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
                  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kCThreads, smem) == cudaSuccess &&
                  per_sm >= 1;
        (void)cudaGetLastError();
        pr.grid = ok ? (sms < kMaxReduceBlocks ? sms : kMaxReduceBlocks) : 0;
        pr.state = ok ? 1 : -1;
        if (!ok) return false;
    }
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    void* args[] = {(void*)&x, (void*)&y, (void*)&numel, (void*)&ws, (void*)&x_min, (void*)&x_max, (void*)&beta,
                    (void*)&beta_t, (void*)&k};
    const cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(pr.grid), dim3(kCThreads), args, smem, st);
    if (e != cudaSuccess) {          // e.g. the grid cannot be co-resident right now: leave it to the other path
        (void)cudaGetLastError();
        return false;
    }
    count_launch();
    *rc = check_launch(tma ? "act_calib_forward(on-chip, TMA)" : "act_calib_forward(on-chip)");
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
        try_onchip_calib(x, y, numel, workspace, x_min, x_max, beta, beta_t, k, (flags & OODFQ_ONCHIP_TMA) != 0, st, &rc))
        return rc;
    rc = launch_minmax(x, numel, workspace, nullptr, x_min, x_max, beta, beta_t, sym ? 1 : 0, st);
    if (rc != OODFQ_OK || !y) return rc;
    return launch_fakequant_scalar(x, y, codes, numel, x_min, x_max, k, sym, /*reverse=*/true, st);
}
