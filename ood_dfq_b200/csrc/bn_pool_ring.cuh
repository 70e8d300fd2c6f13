// Stem forward, TMA-staged: BN -> ReLU -> [QuantAct] -> MaxPool2d(3,2,1), channels_last (included by bn_pool.cu).
//
// The register kernel (bn_pool_fwd_kernel) evaluates every input pixel 1.5 times, keeps a (key, meta, x) triple
// per candidate with three selects per pixel and channel, and hides its own global-load latency in registers:
// ~400 instructions per output vector at 128 registers, 50 % issue-active at 24 % warps active, 63-79 % of the
// HBM rate (profiles/r2_microbench.txt).  Here
//   * a producer warp streams whole input rows (W*C floats, contiguous) into a ring of shared-memory slots with
//     cp.async.bulk; consumer warps never touch global memory for x, so no register is spent on latency;
//   * a consumer thread owns TWO adjacent windows of a row (5 pixel columns instead of 6) and walks down the
//     output rows carrying the horizontal result of input row 2ho+1 into the next window: 5 pixel evaluations per
//     2 outputs and row instead of 12;
//   * with a strictly increasing dequantisation table the candidate is ONE unsigned integer -- the bits of
//     (code + 1.5*2^23) shifted left by four, plus a 4-bit tie-break that orders the window positions in reverse
//     scan order -- so "first maximum in scan order" is eight integer max operations per channel, no selects;
//     a NaN anywhere in a window makes the maximum exceed every regular value and sends that one output vector
//     through a scalar loop with ATen's rule (a later NaN replaces an earlier one);
//   * the ReLU bit and the normalised input are taken from the WINNER only (one shared-memory word per channel,
//     the rows are still in the ring).
// Results are bit-identical to the register kernel (tests/test_gpu_fused.py runs both against ATen's max_pool2d).
#pragma once

namespace oodfq {

constexpr int kPfMaxConsumers = 480;      // + the producer warp = 16 warps, 4 per scheduler: 128 registers each
constexpr int kPfMaxSlots = 8;
constexpr int kPfMinSlots = 5;           // rows 2ho-1, 2ho, 2ho+1 live + two in flight
constexpr int kPfSmemBudget = 208 * 1024;

struct PoolFwdPlan {
    int seg, nseg;        // output rows per item, items per image
    int lanes;            // window PAIRS across a row; consumer threads = lanes * cols rounded up to whole warps
    int consumers;
    int slots;            // ring depth in input rows
    int row_bytes;        // W * C * 4
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

struct RingPos {
    int slot;
    uint32_t phase;
    __device__ __forceinline__ void advance(int K) {
        if (++slot == K) { slot = 0; phase ^= 1u; }
    }
};

enum { KM_PLAIN = 0, KM_PACKED = 1, KM_VALUE = 2 };

// tie-break of window position (r, j): larger = earlier in scan order
//   tb = 4*(2-r) + (2-j);   r = 2 - (tb >> 2), j = 2 - (tb & 3)
constexpr uint32_t kNegInfBits = 0xFF800000u;

template <int KM>
__device__ __forceinline__ uint32_t pixel_key(float x, float a, float b, const QParams& qp, float lowc, const float* lut,
                                              int qh, int qmask) {
    const float zr = fmaf(x, a, b);                                   // BN affine
    if (KM == KM_PACKED) {
        // code of relu(zr) as in bn_pool_fwd_kernel: u = s*z - zp is monotone in z and equals -zp at z = +-0, so
        // max(u, max(-zp, qlo)) folds the ReLU and the lower clamp; (u + 1.5*2^23) rounds to nearest-even and leaves
        // the code in the low mantissa bits.  Shifted left by four the bits still order like the code (the carry out
        // of bit 31 is the same for every regular value) and a NaN becomes >= 0xF8000000.
        float u = __fsub_rn(__fmul_rn(qp.scale, zr), qp.zp);
        u = min_nan(max_nan(u, lowc), qp.qhi);
        return __float_as_uint(__fadd_rn(u, kRoundMagic)) << 4;
    }
    const float z = max_nan(zr, 0.0f);                                 // ReLU (NaN stays NaN)
    if (KM == KM_PLAIN) return __float_as_uint(z);
    const float q = code_of<false>(z, qp);
    const float y = lut[lut_index(q, qh, qmask)];
    return __float_as_uint((q != q) ? q : y);
}

// horizontal result of one input row for this thread's two windows
struct HRow {
    uint32_t k[2][4];     // KM_PACKED: packed key with the column tie-break; else float bits of the key
    int t[2][4];          // float modes: column tie-break (2 - j) of the row's first maximum
};

template <int KM>
__device__ __forceinline__ void hrow_invalid(HRow& h) {
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int c = 0; c < 4; ++c) { h.k[w][c] = KM == KM_PACKED ? 0u : kNegInfBits; h.t[w][c] = 0; }
}

// `row`: this thread's channel vector of pixel column pc0 in the staged row (may be out of the image: never
// dereferenced then); pv[i]: pixel column pc0 + i exists (column pc0 + 1 always does).  ROFF = 4 * (2 - r) for the
// window row r the staged row plays (a row that is carried on is evaluated as r = 2 and gets its +8 when combined).
template <int KM, int ROFF>
__device__ __forceinline__ void hrow_eval(HRow& h, const float4* row, const int cols, const bool (&pv)[5],
                                          const float (&a)[4], const float (&b)[4], const QParams& qp, const float lowc,
                                          const float* lut, const int qh, const int qmask) {
    uint32_t P[5][4];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (i == 1 || pv[i]) {
            const float4 v = row[i * cols];
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) P[i][c] = pixel_key<KM>(xs[c], a[c], b[c], qp, lowc, lut, qh, qmask);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) P[i][c] = KM == KM_PACKED ? 0u : kNegInfBits;
        }
    }
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (KM == KM_PACKED) {
                h.k[w][c] = max(max(P[2 * w][c] + (ROFF + 2u), P[2 * w + 1][c] + (ROFF + 1u)), P[2 * w + 2][c] + (unsigned)ROFF);
                h.t[w][c] = 0;
            } else {
                const float f0 = __uint_as_float(P[2 * w][c]), f1 = __uint_as_float(P[2 * w + 1][c]),
                            f2 = __uint_as_float(P[2 * w + 2][c]);
                const float m = max_nan(max_nan(f0, f1), f2);
                h.k[w][c] = __float_as_uint(m);
                int t = ROFF;                                   // selects, not branches: later candidates first
                t = (f1 == m) ? ROFF + 1 : t;
                t = (f0 == m) ? ROFF + 2 : t;
                h.t[w][c] = t;
            }
        }
}

// One output vector the slow way: ATen's scan (a later element replaces the running maximum when it is greater
// or NaN).  Only entered when a NaN sits in the window.  rows[r]: channel vector of pixel column pcw (window
// column 0) in staged input row r, or nullptr when that row is outside the image.
template <int KM>
__device__ __forceinline__ void slow_window(const float4* const (&rows)[3], const int cols, const bool (&cv)[3],
                                         const float (&a)[4], const float (&b)[4], const QParams& qp, const float lowc,
                                         const float* lut, const int qh, const int qmask, float (&y)[4], int (&tb)[4]) {
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < 4; ++c) tb[c] = 0;
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
        if (!rows[r]) continue;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
            if (!cv[j]) continue;
            const float4 v = rows[r][j * cols];
            const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float key;
                if (KM == KM_PACKED) {
                    const float zr = fmaf(xs[c], a[c], b[c]);
                    float u = __fsub_rn(__fmul_rn(qp.scale, zr), qp.zp);
                    u = min_nan(max_nan(u, lowc), qp.qhi);
                    key = __fadd_rn(u, kRoundMagic);
                } else {
                    key = __uint_as_float(pixel_key<KM>(xs[c], a[c], b[c], qp, lowc, lut, qh, qmask));
                }
                if (takes_over(key, best[c])) { best[c] = key; tb[c] = 4 * (2 - r) + (2 - j); }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (KM == KM_PACKED) {
            const float v = lut[(__float_as_int(best[c]) + qh) & qmask];
            y[c] = (best[c] != best[c]) ? best[c] : v;
        } else {
            y[c] = best[c];
        }
    }
}

template <int KM, bool XHAT>
__device__ __forceinline__ void pool_ring_consume(const unsigned char* __restrict__ ring, uint64_t* full, uint64_t* empty,
                                                  float* __restrict__ out, uint8_t* __restrict__ idx,
                                                  float* __restrict__ xhat, const PoolGeom& G, const PoolFwdPlan& L,
                                                  const BnParams2& P, const QParams& qp, const float* lut, const int qh,
                                                  const int qmask) {
    const int K = L.slots;
    const int col = threadIdx.x % G.cols, lane_w = threadIdx.x / G.cols;
    const bool active = lane_w < L.lanes;
    const int lane = threadIdx.x & 31;
    const int wo0 = 2 * lane_w;                        // this thread's windows: wo0, wo0 + 1
    const int pc0 = 2 * wo0 - 1;                       // pixel columns pc0 .. pc0 + 4
    bool pv[5], wv[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) pv[i] = active && pc0 + i >= 0 && pc0 + i < G.W;
#pragma unroll
    for (int w = 0; w < 2; ++w) wv[w] = active && wo0 + w < G.Wo;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, rm[4] = {0.f, 0.f, 0.f, 0.f}, inv[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { affine2(P, 4 * col + j, a[j], b[j], inv[j]); rm[j] = __ldg(P.rm + 4 * col + j); }
    }
    const float lowc = fmaxf(-qp.zp, qp.qlo);
    // offset (in float4) of this thread's vector of pixel column pc0 inside a staged row; negative for lane 0
    const int voff = pc0 * G.cols + col;
    const long long items = (long long)G.N * L.nseg;
    RingPos cur{0, 0u};
    // shared-memory byte address of this thread's vector of window w's LAST pixel column (j = 2) in slot 0
    const uint32_t cstep = 16u * (uint32_t)G.cols;
    const uint32_t wbase[2] = {smem_u32(ring) + 16u * (uint32_t)(voff + 2 * G.cols), smem_u32(ring) + 16u * (uint32_t)(voff + 4 * G.cols)};
    auto slot_row = [&](const RingPos& p) { return reinterpret_cast<const float4*>(ring + (size_t)p.slot * L.row_bytes) + voff; };
    auto release = [&](const RingPos& p) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[p.slot]);
    };
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const long long n = it / L.nseg;
        const int ho0 = (int)(it % L.nseg) * L.seg, ho1 = min(G.Ho, ho0 + L.seg);
        HRow ca, h1;            // ca: the row carried from the previous window (evaluated as r = 2), h1: row 2ho
        RingPos pa = cur, pb{0, 0u}, pc{0, 0u};
        bool have_a = ho0 > 0;
        const float4 *ra = nullptr, *rb = nullptr, *rc = nullptr;
        hrow_invalid<KM>(ca);
        if (have_a) {
            mbar_wait(&full[pa.slot], pa.phase);
            ra = slot_row(pa);
            if (active) hrow_eval<KM, 0>(ca, ra, G.cols, pv, a, b, qp, lowc, lut, qh, qmask);
            cur.advance(K);
        }
        long long o = ((n * G.Ho + ho0) * G.Wo + wo0) * G.cols + col;
        const int ostride = G.Wo * G.cols;
        for (int ho = ho0; ho < ho1; ++ho, o += ostride) {
            const bool have_c = 2 * ho + 1 < G.H;
            pb = cur; cur.advance(K);
            mbar_wait(&full[pb.slot], pb.phase);
            rb = slot_row(pb);
            if (have_c) { pc = cur; cur.advance(K); }
            if (active) {
                hrow_eval<KM, 4>(h1, rb, G.cols, pv, a, b, qp, lowc, lut, qh, qmask);
                // rows 2ho-1 and 2ho combined; the carried registers are then free for row 2ho+1
                uint32_t pk[2][4];
                int pt[2][4];
#pragma unroll
                for (int w = 0; w < 2; ++w)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (KM == KM_PACKED) {
                            pk[w][c] = max(ca.k[w][c] + 8u, h1.k[w][c]);
                            pt[w][c] = 0;
                        } else {
                            const float k0 = __uint_as_float(ca.k[w][c]), k1 = __uint_as_float(h1.k[w][c]);
                            const float m = max_nan(k0, k1);
                            pk[w][c] = __float_as_uint(m);
                            pt[w][c] = (k0 == m) ? ca.t[w][c] + 8 : h1.t[w][c];
                        }
                    }
                if (have_c) {
                    mbar_wait(&full[pc.slot], pc.phase);
                    rc = slot_row(pc);
                    hrow_eval<KM, 0>(ca, rc, G.cols, pv, a, b, qp, lowc, lut, qh, qmask);
                } else {
                    rc = nullptr;
                    hrow_invalid<KM>(ca);
                }
                const uint32_t sc2 = (uint32_t)(pb.slot == 0 ? K - 1 : pb.slot - 1) + 2u;   // slot of row 2ho-1, plus two
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    if (!wv[w]) continue;
                    float y[4];
                    int tb[4];
                    bool nan = false;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (KM == KM_PACKED) {
                            const uint32_t v = max(pk[w][c], ca.k[w][c]);
                            nan = nan || v >= 0xC0000000u;
                            y[c] = lut[((v >> 4) + qh) & qmask];
                            tb[c] = (int)(v & 15u);
                        } else {
                            const float k01 = __uint_as_float(pk[w][c]), k2 = __uint_as_float(ca.k[w][c]);
                            const float m = max_nan(k01, k2);
                            nan = nan || m != m;
                            y[c] = m;
                            tb[c] = (k01 == m) ? pt[w][c] : ca.t[w][c];
                        }
                    }
                    if (nan) {
                        const float4* const rows[3] = {have_a ? ra + 2 * w * G.cols : nullptr, rb + 2 * w * G.cols,
                                                       have_c ? rc + 2 * w * G.cols : nullptr};
                        const bool cv[3] = {pv[2 * w], pv[2 * w + 1], pv[2 * w + 2]};
                        slow_window<KM>(rows, G.cols, cv, a, b, qp, lowc, lut, qh, qmask, y, tb);
                    }
                    unsigned char code[4];
                    float xh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int rr = tb[c] >> 2, jj = tb[c] & 3;     // 2 - r, 2 - j of the winner
                        bool relu;
                        if (KM == KM_PLAIN && !XHAT) {
                            relu = y[c] > 0.0f;                      // y = max(zr, 0): open exactly where zr > 0
                        } else {
                            // the winner's input: rows 2ho-1, 2ho, 2ho+1 sit in consecutive ring slots (mod K);
                            // t = slot(2ho-1) + r < 2K, and umin(t, t - K) wraps it without a predicate
                            const uint32_t t = sc2 - (uint32_t)rr;
                            const uint32_t slot = min(t, t - (uint32_t)K);
                            const float xw = lds_f32(wbase[w] + slot * (uint32_t)L.row_bytes - (uint32_t)jj * cstep + 4u * c);
                            relu = fmaf(xw, a[c], b[c]) > 0.0f;
                            xh[c] = (xw - rm[c]) * inv[c];
                        }
                        code[c] = (unsigned char)((8 - 3 * rr - jj) | (relu ? 128 : 0));   // window-local index 3r + j
                    }
                    const long long ow = o + (long long)w * G.cols;
                    st_out(reinterpret_cast<float4*>(out) + ow, make_float4(y[0], y[1], y[2], y[3]));
                    reinterpret_cast<uchar4*>(idx)[ow] = make_uchar4(code[0], code[1], code[2], code[3]);
                    if (XHAT) st_out(reinterpret_cast<float4*>(xhat) + ow, make_float4(xh[0], xh[1], xh[2], xh[3]));
                }
            } else if (have_c) {
                mbar_wait(&full[pc.slot], pc.phase);
            }
            // rows 2ho-1 and 2ho are done with; row 2ho+1 stays (carried into the next window, winner lookups)
            if (have_a) release(pa);
            release(pb);
            pa = pc; ra = rc; have_a = have_c;
        }
        if (have_a) release(pa);
    }
}

// producer: one thread walks the same (item, row) sequence and keeps the ring full
__device__ __forceinline__ void pool_ring_produce(unsigned char* ring, uint64_t* full, uint64_t* empty, const float* x,
                                                  const PoolGeom& G, const PoolFwdPlan& L) {
    const int K = L.slots;
    const long long items = (long long)G.N * L.nseg;
    RingPos p{0, 0u};
    long long q = 0;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const long long n = it / L.nseg;
        const int ho0 = (int)(it % L.nseg) * L.seg, ho1 = min(G.Ho, ho0 + L.seg);
        const int first = max(0, 2 * ho0 - 1), last = min(G.H - 1, 2 * ho1 - 1);
        for (int row = first; row <= last; ++row, ++q) {
            if (q >= K) mbar_wait(&empty[p.slot], p.phase ^ 1u);
            mbar_arrive_expect_tx(&full[p.slot], (uint32_t)L.row_bytes);
            bulk_g2s(ring + (size_t)p.slot * L.row_bytes, x + ((n * G.H + row) * (long long)G.W) * G.C,
                     (uint32_t)L.row_bytes, &full[p.slot]);
            p.advance(K);
        }
    }
}

template <bool QUANT, bool XHAT>
__global__ void __launch_bounds__(kPfMaxConsumers + 32, 1)
bn_pool_fwd_tma_kernel(const float* __restrict__ x, float* __restrict__ out, uint8_t* __restrict__ idx,
                       float* __restrict__ xhat, const PoolGeom G, const PoolFwdPlan L, const BnParams2 P,
                       const float* __restrict__ fq_lo, const float* __restrict__ fq_hi, int fq_k) {
    extern __shared__ __align__(128) unsigned char pf_ring[];
    __shared__ __align__(8) uint64_t full[kPfMaxSlots], empty[kPfMaxSlots];
    __shared__ float lut[QUANT ? kLutMax : 1];
    QParams qp = given_qparams(1.0f, 0.0f, 1);
    const int qh = 1 << (fq_k - 1), qmask = (1 << fq_k) - 1;
    bool strict = false;
    if (threadIdx.x == 0) {
        for (int k = 0; k < L.slots; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], (uint32_t)(L.consumers / 32)); }
        mbar_fence_init();
    }
    if (QUANT) {
        qp = make_qparams(__ldg(fq_lo), __ldg(fq_hi), fq_k);
        build_lut(lut, qp, fq_k, threadIdx.x, blockDim.x);
        __syncthreads();
        bool inc = true;
        for (int j = threadIdx.x; j + 1 < (1 << fq_k); j += blockDim.x) inc = inc && (lut[j] < lut[j + 1]);
        strict = __syncthreads_and(inc);
    } else {
        __syncthreads();
    }
    if ((int)threadIdx.x >= L.consumers) {
        if ((int)threadIdx.x == L.consumers) pool_ring_produce(pf_ring, full, empty, x, G, L);
        return;
    }
    if (!QUANT) pool_ring_consume<KM_PLAIN, XHAT>(pf_ring, full, empty, out, idx, xhat, G, L, P, qp, lut, qh, qmask);
    else if (strict) pool_ring_consume<KM_PACKED, XHAT>(pf_ring, full, empty, out, idx, xhat, G, L, P, qp, lut, qh, qmask);
    else pool_ring_consume<KM_VALUE, XHAT>(pf_ring, full, empty, out, idx, xhat, G, L, P, qp, lut, qh, qmask);
}

// plan of the TMA-staged forward, or false when the geometry does not fit (the register kernel takes over)
static bool make_pool_fwd_plan(const PoolGeom& G, PoolFwdPlan& L, size_t& smem, int& threads) {
    const long long row_bytes = (long long)G.W * G.C * 4;
    if (row_bytes > kPfSmemBudget / kPfMinSlots) return false;
    L.row_bytes = (int)row_bytes;
    L.slots = (int)(kPfSmemBudget / row_bytes);
    if (L.slots > kPfMaxSlots) L.slots = kPfMaxSlots;
    L.lanes = (G.Wo + 1) / 2;
    const long long active = (long long)L.lanes * G.cols;
    if (active > kPfMaxConsumers) return false;
    L.consumers = (int)((active + 31) / 32 * 32);
    threads = L.consumers + 32;
    smem = (size_t)L.slots * L.row_bytes;
    // items per image: the split whose last round of CTAs is fullest, counting the halo row a segment re-reads
    double best = -1.0;
    L.nseg = 1;
    L.seg = G.Ho;
    for (int ns = 1; ns <= 16 && ns <= G.Ho; ++ns) {
        const int seg = (G.Ho + ns - 1) / ns;
        const int real = (G.Ho + seg - 1) / seg;
        const double items = (double)G.N * real;
        const double rounds = (double)(long long)((items + kNumSM - 1) / kNumSM);
        const double eff = items / (rounds * kNumSM) * (2.0 * seg / (2.0 * seg + 1.0));
        if (eff > best + 1e-9) { best = eff; L.nseg = real; L.seg = seg; }
    }
    return true;
}

}  // namespace oodfq
