// Shared device helpers for the sm_100a fake-quant kernels.
//
// Arithmetic contract (SURVEY.md section 8(a')): the reference runs every step of
// quantise -> clamp -> dequantise as its own ATen kernel, so each step rounds to
// fp32 separately.  All arithmetic here therefore uses the explicit
// round-to-nearest intrinsics (__fmul_rn / __fsub_rn / __fadd_rn / __fdiv_rn /
// __frcp_rn), which nvcc never contracts into FMAs, and rintf (ties-to-even).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/oodfq_b200.h"

namespace oodfq {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

// Device scratch shared by the reducing kernels (one per stream, zeroed once by the
// caller).  Tickets are self-resetting: the block that draws the last ticket of a
// launch finishes the reduction and puts the counter back to zero.
constexpr int kWsTickets = 64;
constexpr int kMaxBnChannels = 8192;
constexpr int kMaxBnSplit = 64;
constexpr int kMaxReduceBlocks = 4096;
struct Workspace {
    int ticket[kWsTickets];
    int bn_ticket[kMaxBnChannels];
    float mm_partial[2 * kMaxReduceBlocks];
    double bn_partial[2 * kMaxBnSplit * kMaxBnChannels];   // fp64 (S1, S2) per CTA and channel
};
constexpr size_t kWorkspaceBytes = sizeof(Workspace);

// ---- host side bookkeeping (api.cu) -------------------------------------------------
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);
int launch_fakequant_scalar(const float* x, float* y, int8_t* codes, long long numel, const float* lo,
                            const float* hi, int k, bool sym, bool reverse, cudaStream_t st);

// ---- NaN-propagating min / max (torch.min, torch.max and torch.clamp keep NaN) ------
__device__ __forceinline__ float min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// ---- streaming global access ------------------------------------------------------------
// sm_100a has 256-bit global loads/stores (LDG.256 / STG.256); ptxas only accepts the
// direct .L2::evict_first qualifier on that width.  x is read exactly once: no L1
// allocation, evict-first in L2.  y is consumed by the next convolution: default L2 policy.
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld_stream8(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                   "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
// keeps the default L2 policy (the tensor is read again by a following pass)
__device__ __forceinline__ f8 ld_keep8(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                   "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
// coherent load for the in-place case where y aliases x
__device__ __forceinline__ f8 ld_plain8(const float* p) {
    f8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]),
                   "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_out8(float* p, const f8& r) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                    "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_out(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- TMA bulk copies (cp.async.bulk, 1-D) and the mbarrier that counts their bytes ------------------------
// Used where a CTA stages a contiguous run of global memory in shared memory (the single-pass calibrating
// QuantAct, the stem re-layout): one thread issues the copy, the copy engine moves the data without passing
// through registers, every consumer waits on the barrier's phase.  SASS: UBLKCP.S.G, SYNCS.ARRIVE.TRANS64,
// SYNCS.PHASECHK.TRANS64.TRYWAIT.
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
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(phase) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}

// shared -> global bulk copy (size and both addresses multiples of 16 bytes); completion via bulk groups
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N bulk groups of this thread are still READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
// make generic-proxy writes to shared memory visible to the async proxy (before a shared -> global bulk copy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- quantisation parameters ----------------------------------------------------------
struct QParams {
    float scale;  // n / clamp(hi - lo, 1e-8)  evaluated as reciprocal * n
    float zp;     // rint(scale * lo) + 2^(k-1)
    float qlo;    // -2^(k-1)
    float qhi;    //  2^(k-1) - 1
};

// quant_utils.py:117-128.  `n / tensor` in the reference is Tensor.__rdiv__ =
// tensor.reciprocal() * n, hence reciprocal-then-multiply and not a division.
__device__ __forceinline__ QParams make_qparams(float lo, float hi, int k) {
    QParams p;
    float r = __fsub_rn(hi, lo);
    r = max_nan(r, 1e-8f);
    float inv = __frcp_rn(r);
    p.scale = __fmul_rn(inv, (float)((1 << k) - 1));
    float z = rintf(__fmul_rn(p.scale, lo));
    float h = (float)(1 << (k - 1));
    p.zp = __fadd_rn(z, h);
    p.qlo = -h;
    p.qhi = h - 1.0f;
    return p;
}

__device__ __forceinline__ QParams given_qparams(float scale, float zp, int k) {
    QParams p;
    float h = (float)(1 << (k - 1));
    p.scale = scale;
    p.zp = zp;
    p.qlo = -h;
    p.qhi = h - 1.0f;
    return p;
}

// integer code of one element: quant_utils.py:81 (:212 symmetric) then :151-152
template <bool SYM>
__device__ __forceinline__ float code_of(float x, const QParams& p) {
    float a = __fmul_rn(p.scale, x);
    float b = SYM ? a : __fsub_rn(a, p.zp);
    float q = rintf(b);
    return min_nan(max_nan(q, p.qlo), p.qhi);
}

// quant_utils.py:104 (:235 symmetric): true IEEE division
template <bool SYM>
__device__ __forceinline__ float value_of(float q, const QParams& p) {
    float c = SYM ? q : __fadd_rn(q, p.zp);
    return __fdiv_rn(c, p.scale);
}

template <bool SYM>
__device__ __forceinline__ float fake_quant(float x, const QParams& p) {
    return value_of<SYM>(code_of<SYM>(x, p), p);
}

// ---- dequantisation by table ------------------------------------------------------------
// A clamped code takes one of 2^k values, so (q + zp) / scale takes one of 2^k values too.
// The generic IEEE division is the expensive part of the element loop (MUFU.RCP + FCHK +
// 5 FFMA, and its slow path is entered whenever the numerator is zero, i.e. for every
// post-ReLU zero).  For k <= 8 each CTA therefore evaluates the 2^k quotients once with the
// true division -- bit-exact by construction -- and the loop does a conflict-free shared
// memory lookup instead.
constexpr int kLutMax = 256;
constexpr float kRoundMagic = 12582912.0f;   // 1.5 * 2^23: float(kRoundMagic + n) has bits 0x4B400000 + n

__device__ __forceinline__ void build_lut(float* lut, const QParams& p, int k, int tid, int nthreads) {
    for (int j = tid; j < (1 << k); j += nthreads) lut[j] = value_of<false>(__fadd_rn((float)j, p.qlo), p);
}

// index of a clamped, non-NaN code q in [-h, h-1] (h = -qlo): low bits of the magic-number sum
__device__ __forceinline__ int lut_index(float q, int h, int mask) {
    return (__float_as_int(__fadd_rn(q, kRoundMagic)) + h) & mask;
}

// The code only ever serves as the table index, so its rounding and clamping are done the cheap way round: clamp
// and rint commute when the bounds are integers, and for |u| <= 2^22 the sum (u + 1.5*2^23) IS rint(u) (ties to
// even) sitting in the low mantissa bits -- no FRND (a quarter-rate conversion-pipe instruction), one FADD fewer.
__device__ __forceinline__ float fake_quant_lut(float x, const QParams& p, const float* lut, int h, int mask) {
    float u = __fsub_rn(__fmul_rn(p.scale, x), p.zp);
    u = min_nan(max_nan(u, p.qlo), p.qhi);
    const float key = __fadd_rn(u, kRoundMagic);
    const float y = lut[(__float_as_int(key) + h) & mask];
    return (key != key) ? key : y;           // NaN in, NaN out (as the ATen chain does)
}

// fakequant(relu(z)) with the ReLU folded into the lower clamp.  u(z) = scale*z - zp (two roundings) is monotone
// in z and u(+-0) = -zp exactly, so u(relu(z)) = max(u(z), -zp); zp is an integer (make_qparams rounds it), so that
// max commutes with rint like the clamp does: code = min(max(rint(u), lowc), qhi) with lowc = max(-zp, qlo).
// Bit-identical to relu -> fake_quant_lut (NaN stays NaN), three instructions shorter.  NOT valid for a zero-point
// handed in from outside (OODFQ_PARAMS_GIVEN), which need not be an integer.
__device__ __forceinline__ float relu_lower_bound(const QParams& p) { return fmaxf(-p.zp, p.qlo); }
__device__ __forceinline__ float relu_fake_quant_lut(float z, const QParams& p, float lowc, const float* lut, int h,
                                                     int mask) {
    float u = __fsub_rn(__fmul_rn(p.scale, z), p.zp);
    u = min_nan(max_nan(u, lowc), p.qhi);
    const float key = __fadd_rn(u, kRoundMagic);
    const float y = lut[(__float_as_int(key) + h) & mask];
    return (key != key) ? key : y;
}

template <int MODE, bool SYM>
__device__ __forceinline__ float apply_mode(float x, const QParams& p) {
    if (MODE == OODFQ_MODE_FAKEQUANT) return fake_quant<SYM>(x, p);
    if (MODE == OODFQ_MODE_QUANTIZE) {
        float a = __fmul_rn(p.scale, x);
        return rintf(SYM ? a : __fsub_rn(a, p.zp));
    }
    return value_of<SYM>(x, p);  // DEQUANTIZE
}

// ---- warp / block reductions ------------------------------------------------------------
__device__ __forceinline__ float warp_min_nan(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_nan(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Bias-corrected running range, quant_modules.py:87-89.  The corrected value is stored
// back into the state.  beta_t_new = beta_t * beta has been formed by the caller.
__device__ __forceinline__ float ema_step(float state, float sample, float beta, float beta_t_new) {
    float omb = __fsub_rn(1.0f, beta);
    float t1 = __fmul_rn(state, beta);
    float t2 = __fmul_rn(sample, omb);
    float t3 = __fadd_rn(t1, t2);
    float d = __fsub_rn(1.0f, beta_t_new);
    return __fdiv_rn(t3, d);
}

// resident CTAs per SM of a kernel at a block size (persistent grids are sized with it);
// answered by the occupancy calculator from the cubin, once per kernel instantiation
template <typename K>
inline int resident_ctas(K kernel, int threads, size_t dyn_smem = 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = 4;
    }
    return n;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

}  // namespace oodfq
