// Per-output-channel weight fake-quantisation for MANY layers in one launch.
//
// Replaces, for every Quant_Conv2d / Quant_Linear of a model at once, the
// per-layer chain in quantization_utils/quant_modules.py:266-279 (:215-230,
// DSG :420-431, :465-479): view(C_out,-1) -> min(dim=1) -> max(dim=1) -> the six
// fake-quant passes (about 16 launches per layer, index tensors materialised).
// Here each output row is read from HBM once, reduced on chip, and its
// fake-quantised twin written once: 8 B/elem, one launch for the whole model.
//
// Work split: rows of <= kWarpRowMax elements are handled one per warp (8 rows per CTA, no
// block barrier, one dequantisation table per warp), longer rows get a whole CTA; either way
// the row is reduced, then re-read through L1 (a 4608-element row is 18 KB) and quantised.
//
// Roofline: HBM, 8 algorithmic bytes per weight element.
#include "common.cuh"

namespace oodfq {

constexpr int kWThreads = 256;
constexpr int kWarpsPerCta = kWThreads / 32;
constexpr int kWarpRowMax = 8192;   // 32 KB: a row still sits in L1 between the two passes of its warp
constexpr int kMaxJobs = 40;        // descriptors travel in kernel-parameter space (< 4 KB)

struct WeightJob {
    const float* w;
    float* wq;
    float* lo;
    float* hi;
    int8_t* codes;
    long long rows;
    long long row_len;
    int k;
    int flags;
    int first_block;   // first CTA of this job in the fused grid
    int rows_per_cta;  // kWarpsPerCta (warp-per-row) or 1 (CTA-per-row)
};

struct WeightBatch {
    WeightJob job[kMaxJobs];
    int n;
};

__device__ __forceinline__ void range_from(float mn, float mx, bool sym, float& lo, float& hi) {
    if (sym) {   // quant_modules.py:473-474: +-max|w|  (mn, mx hold min / max of |w| here)
        lo = -mx;
        hi = mx;
    } else {
        lo = mn;
        hi = mx;
    }
}

// One warp per row (row_len <= kWarpRowMax = 4 KB): pass 1 reduces the row, pass 2 re-reads it
// through L1 (the line was just brought in by this very warp) and quantises.  Keeping the row
// in L1 instead of 32 registers per lane is what lets 6-8 CTAs stay resident per SM.
template <bool SYM>
__device__ __forceinline__ void warp_row(const WeightJob& jb, long long row, int lane, float* lut) {
    const int len = (int)jb.row_len;
    const float* src = jb.w + row * jb.row_len;
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    const bool vec = ((len & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(jb.wq + row * jb.row_len) & 15u) == 0);
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
        for (int i = lane; i < (len >> 2); i += 32) {
            float4 t = __ldg(s4 + i);
            if (SYM) { t.x = fabsf(t.x); t.y = fabsf(t.y); t.z = fabsf(t.z); t.w = fabsf(t.w); }
            mn = min_nan(min_nan(mn, t.x), min_nan(t.y, min_nan(t.z, t.w)));
            mx = max_nan(max_nan(mx, t.x), max_nan(t.y, max_nan(t.z, t.w)));
        }
    } else {
#pragma unroll 4
        for (int i = lane; i < len; i += 32) {
            float t = __ldg(src + i);
            if (SYM) t = fabsf(t);
            mn = min_nan(mn, t);
            mx = max_nan(mx, t);
        }
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    float lo, hi;
    range_from(mn, mx, SYM, lo, hi);
    const QParams p = make_qparams(lo, hi, jb.k);
    if (lane == 0) {
        if (jb.lo) jb.lo[row] = lo;
        if (jb.hi) jb.hi[row] = hi;
    }
    float* dst = jb.wq + row * jb.row_len;
    int8_t* cd = jb.codes ? jb.codes + row * jb.row_len : nullptr;
    if (!SYM && !cd && jb.k <= 8) {          // table of the 2^k dequantised values of this row
        build_lut(lut, p, jb.k, lane, 32);
        __syncwarp();
        const int h = 1 << (jb.k - 1), mask = (1 << jb.k) - 1;
        if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
            for (int i = lane; i < (len >> 2); i += 32) {
                float4 t = __ldg(s4 + i);
                d4[i] = make_float4(fake_quant_lut(t.x, p, lut, h, mask), fake_quant_lut(t.y, p, lut, h, mask),
                                    fake_quant_lut(t.z, p, lut, h, mask), fake_quant_lut(t.w, p, lut, h, mask));
            }
        } else {
#pragma unroll 4
            for (int i = lane; i < len; i += 32) dst[i] = fake_quant_lut(__ldg(src + i), p, lut, h, mask);
        }
        return;
    }
    for (int i = lane; i < len; i += 32) {
        float q = code_of<SYM>(__ldg(src + i), p);
        if (cd) cd[i] = (int8_t)q;
        dst[i] = value_of<SYM>(q, p);
    }
}

template <bool SYM>
__device__ __forceinline__ void cta_row(const WeightJob& jb, long long row, float* s_red, float* lut) {
    const long long len = jb.row_len;
    const float* src = jb.w + row * len;
    float* dst = jb.wq + row * len;
    const bool vec = ((len & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
    float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (long long i = threadIdx.x; i < (len >> 2); i += kWThreads) {
            float4 t = __ldg(s4 + i);   // allocate in L1: read again below
            if (SYM) { t.x = fabsf(t.x); t.y = fabsf(t.y); t.z = fabsf(t.z); t.w = fabsf(t.w); }
            mn = min_nan(min_nan(mn, t.x), min_nan(t.y, min_nan(t.z, t.w)));
            mx = max_nan(max_nan(mx, t.x), max_nan(t.y, max_nan(t.z, t.w)));
        }
    } else {
        for (long long i = threadIdx.x; i < len; i += kWThreads) {
            float t = __ldg(src + i);
            if (SYM) t = fabsf(t);
            mn = min_nan(mn, t);
            mx = max_nan(mx, t);
        }
    }
    mn = warp_min_nan(mn);
    mx = warp_max_nan(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp] = mn; s_red[kWarpsPerCta + warp] = mx; }
    __syncthreads();
    mn = s_red[0];
    mx = s_red[kWarpsPerCta];
#pragma unroll
    for (int w = 1; w < kWarpsPerCta; ++w) {
        mn = min_nan(mn, s_red[w]);
        mx = max_nan(mx, s_red[kWarpsPerCta + w]);
    }
    float lo, hi;
    range_from(mn, mx, SYM, lo, hi);
    const QParams p = make_qparams(lo, hi, jb.k);
    if (threadIdx.x == 0) {
        if (jb.lo) jb.lo[row] = lo;
        if (jb.hi) jb.hi[row] = hi;
    }
    int8_t* cd = jb.codes ? jb.codes + row * len : nullptr;
    const bool use_lut = !SYM && !cd && jb.k <= 8;
    const int h = 1 << (jb.k - 1), mask = (1 << jb.k) - 1;
    if (use_lut) {
        build_lut(lut, p, jb.k, threadIdx.x, kWThreads);
        __syncthreads();
    }
    if (vec && !cd) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (long long i = threadIdx.x; i < (len >> 2); i += kWThreads) {
            float4 t = __ldg(s4 + i);
            if (use_lut)
                d4[i] = make_float4(fake_quant_lut(t.x, p, lut, h, mask), fake_quant_lut(t.y, p, lut, h, mask),
                                    fake_quant_lut(t.z, p, lut, h, mask), fake_quant_lut(t.w, p, lut, h, mask));
            else
                d4[i] = make_float4(fake_quant<SYM>(t.x, p), fake_quant<SYM>(t.y, p),
                                    fake_quant<SYM>(t.z, p), fake_quant<SYM>(t.w, p));
        }
    } else if (use_lut) {
        for (long long i = threadIdx.x; i < len; i += kWThreads) dst[i] = fake_quant_lut(__ldg(src + i), p, lut, h, mask);
    } else {
        for (long long i = threadIdx.x; i < len; i += kWThreads) {
            float q = code_of<SYM>(__ldg(src + i), p);
            if (cd) cd[i] = (int8_t)q;
            dst[i] = value_of<SYM>(q, p);
        }
    }
}

__global__ void __launch_bounds__(kWThreads)
weight_fq_kernel(const __grid_constant__ WeightBatch batch) {
    __shared__ float s_red[2 * kWarpsPerCta];
    __shared__ float s_lut[kWarpsPerCta][kLutMax];
    // which job owns this CTA (jobs are few: linear scan over uniform parameter memory)
    int j = 0;
    while (j + 1 < batch.n && (int)blockIdx.x >= batch.job[j + 1].first_block) ++j;
    const WeightJob& jb = batch.job[j];
    const int local = (int)blockIdx.x - jb.first_block;
    const bool sym = (jb.flags & OODFQ_SYMMETRIC) != 0;
    if (jb.rows_per_cta == 1) {
        if (sym) cta_row<true>(jb, local, s_red, s_lut[0]); else cta_row<false>(jb, local, s_red, s_lut[0]);
    } else {
        const long long row = (long long)local * kWarpsPerCta + (threadIdx.x >> 5);
        if (row < jb.rows) {
            float* lut = s_lut[threadIdx.x >> 5];
            if (sym) warp_row<true>(jb, row, threadIdx.x & 31, lut); else warp_row<false>(jb, row, threadIdx.x & 31, lut);
        }
    }
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_weight_fq_multi(const oodfq_weight_desc* d, int n_tensors, oodfq_stream_t stream) {
    if (n_tensors < 0 || (n_tensors > 0 && !d)) return fail(OODFQ_EINVAL, "weight_fq_multi: bad descriptor table");
    cudaStream_t st = (cudaStream_t)stream;
    int done = 0;
    while (done < n_tensors) {
        WeightBatch batch;
        batch.n = 0;
        long long blocks = 0;
        while (done < n_tensors && batch.n < kMaxJobs) {
            const oodfq_weight_desc& s = d[done];
            if (s.rows < 0 || s.row_len < 0) return fail(OODFQ_EINVAL, "weight_fq_multi: tensor %d has a negative shape", done);
            if (s.rows == 0 || s.row_len == 0) { ++done; continue; }
            if (!s.w || !s.wq) return fail(OODFQ_EINVAL, "weight_fq_multi: tensor %d has a null pointer", done);
            if (s.k < 1 || s.k > 16) return fail(OODFQ_EINVAL, "weight_fq_multi: tensor %d k=%d outside [1,16]", done, s.k);
            if (s.codes && s.k > 8) return fail(OODFQ_EINVAL, "weight_fq_multi: tensor %d int8 codes need k <= 8", done);
            WeightJob& jb = batch.job[batch.n];
            jb.w = s.w; jb.wq = s.wq; jb.lo = s.lo; jb.hi = s.hi; jb.codes = s.codes;
            jb.rows = s.rows; jb.row_len = s.row_len; jb.k = s.k; jb.flags = s.flags;
            jb.rows_per_cta = (s.row_len <= kWarpRowMax) ? kWarpsPerCta : 1;
            long long need = (s.rows + jb.rows_per_cta - 1) / jb.rows_per_cta;
            if (blocks + need > 0x7fffffffLL) {
                if (batch.n == 0) return fail(OODFQ_EINVAL, "weight_fq_multi: tensor %d has too many rows", done);
                break;
            }
            jb.first_block = (int)blocks;
            blocks += need;
            ++batch.n;
            ++done;
        }
        if (batch.n == 0) break;
        weight_fq_kernel<<<(unsigned)blocks, kWThreads, 0, st>>>(batch);
        count_launch();
        int rc = check_launch("weight_fq_multi");
        if (rc != OODFQ_OK) return rc;
    }
    return OODFQ_OK;
}
