// Element-wise fake-quantisation (frozen range): one HBM read, one HBM write.
//
// Replaces the six ATen passes of AsymmetricQuantFunction.forward
// (quantization_utils/quant_utils.py:138-157: mul, sub, round, clamp, add, div =
// 48 B/elem) with a single 8 B/elem streaming kernel.  The range arrives as device
// pointers (QuantAct buffers x_min / x_max, quant_modules.py:51-52) and scale /
// zero-point are derived in-kernel, so a forward costs one launch and no host sync.
//
// Roofline: HBM bandwidth, 8 algorithmic bytes per element.
#include "common.cuh"

namespace oodfq {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;  // 4 x 256-bit loads in flight per thread = 32 KB per CTA tile

// LUT = true: asymmetric FAKEQUANT with k <= 8, dequantised values come from a 2^k-entry
// shared-memory table (see common.cuh); otherwise the generic per-element division.
template <int MODE, bool SYM, bool CODES, bool LUT>
__global__ void __launch_bounds__(kThreads)
fq_flat_kernel(const float* x, float* y, int8_t* __restrict__ codes,
               long long numel, const float* __restrict__ p0, const float* __restrict__ p1,
               int k, int given, int aliased, int reverse, int relu) {
    __shared__ float lut[LUT ? kLutMax : 1];
    const QParams p = given ? given_qparams(__ldg(p0), __ldg(p1), k)
                            : make_qparams(__ldg(p0), __ldg(p1), k);
    const int h = 1 << (k - 1), mask = (1 << k) - 1;
    if (LUT) {
        build_lut(lut, p, k, threadIdx.x, kThreads);
        __syncthreads();
    }
    const long long n8 = numel >> 3;                    // 256-bit vectors
    const long long tile = (long long)kThreads * kUnroll;
    const long long ntiles = (n8 + tile - 1) / tile;
    // reverse = 1: walk the tensor back to front, so that a pass which follows a
    // front-to-back read of the same tensor (calibration) starts on its L2-resident tail
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long base = (reverse ? ntiles - 1 - t : t) * tile;
        f8 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            long long i = base + u * kThreads + threadIdx.x;
            if (i < n8) v[u] = aliased ? ld_plain8(x + 8 * i) : ld_stream8(x + 8 * i);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            long long i = base + u * kThreads + threadIdx.x;
            if (i < n8) {
                f8 r;
                if (relu) {                      // the ReLU in front of the QuantAct, NaN-preserving like clamp_min
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[u].v[j] = max_nan(v[u].v[j], 0.0f);
                }
                if (CODES) {
                    union { signed char c[8]; int2 w; } pk;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float q = code_of<SYM>(v[u].v[j], p);
                        pk.c[j] = (signed char)q;
                        r.v[j] = value_of<SYM>(q, p);
                    }
                    reinterpret_cast<int2*>(codes)[i] = pk.w;
                } else if (LUT) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r.v[j] = fake_quant_lut(v[u].v[j], p, lut, h, mask);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r.v[j] = apply_mode<MODE, SYM>(v[u].v[j], p);
                }
                st_out8(y + 8 * i, r);
            }
        }
    }
    // ragged tail (numel % 8 elements)
    if (blockIdx.x == 0) {
        long long i = (n8 << 3) + threadIdx.x;
        if (i < numel) {
            float xv = relu ? max_nan(x[i], 0.0f) : x[i];
            if (CODES) {
                float q = code_of<SYM>(xv, p);
                codes[i] = (int8_t)q;
                y[i] = value_of<SYM>(q, p);
            } else {
                y[i] = apply_mode<MODE, SYM>(xv, p);
            }
        }
    }
}

// Scalar-access twin for buffers that are not 16-byte aligned (views into storage).
template <int MODE, bool SYM, bool CODES>
__global__ void __launch_bounds__(kThreads)
fq_flat_scalar_kernel(const float* x, float* y, int8_t* codes, long long numel,
                      const float* __restrict__ p0, const float* __restrict__ p1, int k, int given, int relu) {
    const QParams p = given ? given_qparams(__ldg(p0), __ldg(p1), k)
                            : make_qparams(__ldg(p0), __ldg(p1), k);
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < numel;
         i += (long long)gridDim.x * kThreads) {
        float xv = relu ? max_nan(x[i], 0.0f) : x[i];
        if (CODES) {
            float q = code_of<SYM>(xv, p);
            codes[i] = (int8_t)q;
            y[i] = value_of<SYM>(q, p);
        } else {
            y[i] = apply_mode<MODE, SYM>(xv, p);
        }
    }
}

// One range per leading-dimension row (the Function called with [C_out] bounds,
// quant_utils.py:70-76).  blockIdx.x = row * chunks + chunk.
template <int MODE, bool SYM, bool CODES>
__global__ void __launch_bounds__(kThreads)
fq_rows_kernel(const float* x, float* y, int8_t* codes, long long row_len,
               const float* __restrict__ p0, const float* __restrict__ p1, int k, int given, int chunks) {
    const long long row = blockIdx.x / chunks;
    const int chunk = blockIdx.x % chunks;
    const QParams p = given ? given_qparams(__ldg(p0 + row), __ldg(p1 + row), k)
                            : make_qparams(__ldg(p0 + row), __ldg(p1 + row), k);
    const float* xr = x + row * row_len;
    float* yr = y + row * row_len;
    int8_t* cr = CODES ? codes + row * row_len : nullptr;
    for (long long i = (long long)chunk * kThreads + threadIdx.x; i < row_len; i += (long long)chunks * kThreads) {
        float xv = xr[i];
        if (CODES) {
            float q = code_of<SYM>(xv, p);
            cr[i] = (int8_t)q;
            yr[i] = value_of<SYM>(q, p);
        } else {
            yr[i] = apply_mode<MODE, SYM>(xv, p);
        }
    }
}

__global__ void quant_params_kernel(const float* __restrict__ lo, const float* __restrict__ hi,
                                    float* scale, float* zp, long long n, int k) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        QParams p = make_qparams(lo[i], hi[i], k);
        scale[i] = p.scale;
        zp[i] = p.zp;
    }
}

template <int MODE, bool SYM, bool CODES>
static int launch_fq(const float* x, float* y, int8_t* codes, long long numel, const float* p0,
                     const float* p1, long long rows, int k, int given, cudaStream_t st, int reverse = 0,
                     int relu = 0) {
    if (rows == 1) {
        const bool vec = aligned32(x) && aligned32(y) && (!CODES || (reinterpret_cast<uintptr_t>(codes) & 7u) == 0);
        if (vec) {
            long long n8 = numel >> 3;
            long long tiles = (n8 + (long long)kThreads * kUnroll - 1) / ((long long)kThreads * kUnroll);
            const int aliased = (const void*)x == (const void*)y;
            constexpr bool kCanLut = (MODE == OODFQ_MODE_FAKEQUANT) && !SYM && !CODES;
            if (kCanLut && k <= 8) {
                static const int per_sm = resident_ctas(fq_flat_kernel<MODE, SYM, CODES, kCanLut>, kThreads);
                long long cap = (long long)kNumSM * per_sm;
                int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
                fq_flat_kernel<MODE, SYM, CODES, kCanLut><<<grid, kThreads, 0, st>>>(x, y, codes, numel, p0, p1, k,
                                                                                    given, aliased, reverse, relu);
            } else {
                static const int per_sm = resident_ctas(fq_flat_kernel<MODE, SYM, CODES, false>, kThreads);
                long long cap = (long long)kNumSM * per_sm;
                int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
                fq_flat_kernel<MODE, SYM, CODES, false><<<grid, kThreads, 0, st>>>(x, y, codes, numel, p0, p1, k,
                                                                                  given, aliased, reverse, relu);
            }
        } else {
            long long blocks = (numel + kThreads - 1) / kThreads;
            long long cap = (long long)kNumSM * 8;
            int grid = (int)(blocks < cap ? blocks : cap);
            fq_flat_scalar_kernel<MODE, SYM, CODES><<<grid, kThreads, 0, st>>>(x, y, codes, numel, p0, p1, k, given, relu);
        }
    } else {
        long long row_len = numel / rows;
        long long want = (row_len + (long long)kThreads * 4 - 1) / ((long long)kThreads * 4);
        int chunks = (int)(want < 1 ? 1 : (want > 64 ? 64 : want));
        long long grid = rows * chunks;
        if (grid > 0x7fffffffLL) return fail(OODFQ_EINVAL, "fq_forward: too many rows (%lld)", rows);
        fq_rows_kernel<MODE, SYM, CODES><<<(unsigned)grid, kThreads, 0, st>>>(x, y, codes, row_len, p0, p1, k, given, chunks);
    }
    count_launch();
    return check_launch("fq_forward");
}

// used by the calibrating path (fq_calib.cu): scalar range, back-to-front
int launch_fakequant_scalar(const float* x, float* y, int8_t* codes, long long numel, const float* lo,
                            const float* hi, int k, bool sym, bool reverse, cudaStream_t st) {
    if (codes) {
        return sym ? launch_fq<OODFQ_MODE_FAKEQUANT, true, true>(x, y, codes, numel, lo, hi, 1, k, 0, st, reverse)
                   : launch_fq<OODFQ_MODE_FAKEQUANT, false, true>(x, y, codes, numel, lo, hi, 1, k, 0, st, reverse);
    }
    return sym ? launch_fq<OODFQ_MODE_FAKEQUANT, true, false>(x, y, codes, numel, lo, hi, 1, k, 0, st, reverse)
               : launch_fq<OODFQ_MODE_FAKEQUANT, false, false>(x, y, codes, numel, lo, hi, 1, k, 0, st, reverse);
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_quant_params(const float* lo, const float* hi, float* scale, float* zero_point,
                                  long long n, int k, oodfq_stream_t stream) {
    if (!lo || !hi || !scale || !zero_point) return fail(OODFQ_EINVAL, "quant_params: null pointer");
    if (k < 1 || k > 16) return fail(OODFQ_EINVAL, "quant_params: k=%d outside [1,16]", k);
    if (n <= 0) return OODFQ_OK;
    quant_params_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(lo, hi, scale, zero_point, n, k);
    count_launch();
    return check_launch("quant_params");
}

extern "C" int oodfq_fq_forward(const float* x, float* y, int8_t* codes, long long numel,
                                const float* p0, const float* p1, long long rows, int k, int mode,
                                int flags, oodfq_stream_t stream) {
    if (numel < 0 || rows < 1) return fail(OODFQ_EINVAL, "fq_forward: numel=%lld rows=%lld", numel, rows);
    if (numel == 0) return OODFQ_OK;
    if (!x || !y || !p0 || !p1) return fail(OODFQ_EINVAL, "fq_forward: null pointer");
    if (k < 1 || k > 16) return fail(OODFQ_EINVAL, "fq_forward: k=%d outside [1,16]", k);
    if (numel % rows != 0) return fail(OODFQ_EINVAL, "fq_forward: numel %lld not divisible by rows %lld", numel, rows);
    if (codes && (k > 8 || mode != OODFQ_MODE_FAKEQUANT))
        return fail(OODFQ_EINVAL, "fq_forward: int8 codes need k <= 8 and FAKEQUANT mode");
    const bool sym = (flags & OODFQ_SYMMETRIC) != 0;
    const int given = (flags & OODFQ_PARAMS_GIVEN) ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int relu = (flags & OODFQ_RELU_FIRST) ? 1 : 0;
    if (relu && (rows != 1 || mode != OODFQ_MODE_FAKEQUANT))
        return fail(OODFQ_EINVAL, "fq_forward: RELU_FIRST needs a scalar range and FAKEQUANT mode");
#define OODFQ_GO(MODE, SYM, CODES) return launch_fq<MODE, SYM, CODES>(x, y, codes, numel, p0, p1, rows, k, given, st, 0, relu)
    switch (mode) {
        case OODFQ_MODE_FAKEQUANT:
            if (codes) { if (sym) OODFQ_GO(OODFQ_MODE_FAKEQUANT, true, true); else OODFQ_GO(OODFQ_MODE_FAKEQUANT, false, true); }
            if (sym) OODFQ_GO(OODFQ_MODE_FAKEQUANT, true, false); else OODFQ_GO(OODFQ_MODE_FAKEQUANT, false, false);
        case OODFQ_MODE_QUANTIZE:
            if (sym) OODFQ_GO(OODFQ_MODE_QUANTIZE, true, false); else OODFQ_GO(OODFQ_MODE_QUANTIZE, false, false);
        case OODFQ_MODE_DEQUANTIZE:
            if (sym) OODFQ_GO(OODFQ_MODE_DEQUANTIZE, true, false); else OODFQ_GO(OODFQ_MODE_DEQUANTIZE, false, false);
        default:
            return fail(OODFQ_EINVAL, "fq_forward: unknown mode %d", mode);
    }
#undef OODFQ_GO
}
