/*
 * fq_oracle.c -- plain-C restatement of the reference fake-quant arithmetic.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): a third, torch-free statement of the
 * algorithm, checked against tests/golden/ by tests/test_oracle_golden.py.  Paths are
 * relative to /root/reference.  Build: oracle/build_c.py (gcc -O2 -ffp-contract=off:
 * every fp32 operation below must stay a separate IEEE operation, exactly like the
 * reference's one-ATen-kernel-per-step execution).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* torch.clamp / min / max propagate NaN; fminf / fmaxf do not */
static float max_nan(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
static float min_nan(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }

/* quantization_utils/quant_utils.py:117-128 (and the DSG twin :248-259).
 * `n / tensor` in the reference is tensor.reciprocal() * n. */
void fqc_params(int k, float lo, float hi, float* scale, float* zp) {
    float r = hi - lo;
    r = max_nan(r, 1e-8f);
    float inv = 1.0f / r;
    float s = inv * (float)((1 << k) - 1);
    float z = s * lo;
    z = rintf(z);                       /* round half to even (default rounding mode) */
    *scale = s;
    *zp = z + (float)(1 << (k - 1));
}

/* quant_utils.py:81 (:212 symmetric) then the clamp of :151-152 */
static float code_of(float x, float scale, float zp, int k, int symmetric) {
    float a = scale * x;
    float b = symmetric ? a : a - zp;
    float q = rintf(b);
    float h = (float)(1 << (k - 1));
    return min_nan(max_nan(q, -h), h - 1.0f);
}

/* quant_utils.py:104 (:235 symmetric): true division */
static float value_of(float q, float scale, float zp, int symmetric) {
    float c = symmetric ? q : q + zp;
    return c / scale;
}

/* AsymmetricQuantFunction.forward / SymmetricQuantFunction_DSG.forward (quant_utils.py:138-157, :268-286)
 * rows == 1: one range; else one (lo[r], hi[r]) per leading row of n/rows elements.
 * codes may be NULL. */
void fqc_fake_quant(const float* x, float* y, float* codes, size_t n, size_t rows, int k,
                    const float* lo, const float* hi, int symmetric) {
    size_t row_len = n / rows;
    for (size_t r = 0; r < rows; ++r) {
        float scale, zp;
        fqc_params(k, lo[r], hi[r], &scale, &zp);
        for (size_t i = r * row_len; i < (r + 1) * row_len; ++i) {
            float q = code_of(x[i], scale, zp, k, symmetric);
            if (codes) codes[i] = q;
            y[i] = value_of(q, scale, zp, symmetric);
        }
    }
}

/* x.data.min(), x.data.max()  (quant_modules.py:81-82) */
void fqc_minmax(const float* x, size_t n, float* mn, float* mx) {
    float a = INFINITY, b = -INFINITY;
    for (size_t i = 0; i < n; ++i) { a = min_nan(a, x[i]); b = max_nan(b, x[i]); }
    *mn = a;
    *mx = b;
}

/* per-row ranges of a weight matrix: quant_modules.py:271-273 (symmetric :473-474) */
void fqc_row_ranges(const float* w, size_t rows, size_t row_len, int symmetric, float* lo, float* hi) {
    for (size_t r = 0; r < rows; ++r) {
        float a = INFINITY, b = -INFINITY;
        for (size_t i = 0; i < row_len; ++i) {
            float t = w[r * row_len + i];
            if (symmetric) t = fabsf(t);
            a = min_nan(a, t);
            b = max_nan(b, t);
        }
        lo[r] = symmetric ? -b : a;
        hi[r] = b;
    }
}

/* One calibrating step of QuantAct: quant_modules.py:87-89 (DSG bounds :369-374).
 * state = {x_min, x_max, beta_t}, updated in place; the corrected value is stored back. */
void fqc_range_update(float* state, float beta, float data_min, float data_max, int symmetric) {
    if (symmetric) {
        float m = max_nan(fabsf(data_min), fabsf(data_max));
        data_min = -m;
        data_max = m;
    }
    float beta_t = state[2] * beta;
    float omb = 1.0f - beta;
    float d = 1.0f - beta_t;
    float t;
    t = state[0] * beta; t = t + data_min * omb; state[0] = t / d;
    t = state[1] * beta; t = t + data_max * omb; state[1] = t / d;
    state[2] = beta_t;
}

/* per-channel batch mean and biased variance of an NCHW tensor (trainer_direct.py:388-393), in fp64 */
void fqc_channel_stats(const float* x, size_t N, size_t C, size_t HW, double* mean, double* var) {
    for (size_t c = 0; c < C; ++c) {
        double s = 0.0, ss = 0.0;
        for (size_t n = 0; n < N; ++n)
            for (size_t i = 0; i < HW; ++i) s += (double)x[(n * C + c) * HW + i];
        double m = s / (double)(N * HW);
        for (size_t n = 0; n < N; ++n)
            for (size_t i = 0; i < HW; ++i) {
                double d = (double)x[(n * C + c) * HW + i] - m;
                ss += d * d;
            }
        mean[c] = m;
        var[c] = ss / (double)(N * HW);
    }
}

/* QuantAct_MSE range search: quant_modules.py:160-178 with find_MSESmallest (quant_utils.py:36-47) and lp_loss
 * (quant_utils.py:26-33, reduction='all').  Candidate i uses the range data_{min,max} * (float)(1.0 - i*0.01);
 * score_i = mean |x - fakequant_i(x)|^p accumulated in fp64 (ATen's fp32 pairwise mean agrees to rounding);
 * the first strict minimum below 1e10 is kept.  state = {x_min, x_max, beta_t}: plain EMA, no bias correction.
 * scores (nullable) receives the `steps` scores; returns the kept candidate (-1: none). */
int fqc_mse_search(const float* x, size_t n, int k, int steps, double p, float beta, float* state, float* scores) {
    float mn, mx;
    fqc_minmax(x, n, &mn, &mx);
    float best = 1e+10f;
    int keep = -1;
    for (int i = 0; i < steps; ++i) {
        const float f = (float)(1.0 - (double)i * 0.01);
        const float lo = mn * f, hi = mx * f;
        float scale, zp;
        fqc_params(k, lo, hi, &scale, &zp);
        double acc = 0.0;
        for (size_t j = 0; j < n; ++j) {
            const float y = value_of(code_of(x[j], scale, zp, k, 0), scale, zp, 0);
            const float d = fabsf(x[j] - y);
            acc += (double)powf(d, (float)p);
        }
        const float score = (float)(acc / (double)n);
        if (scores) scores[i] = score;
        if (score < best) { best = score; keep = i; }
    }
    const int c = keep < 0 ? 0 : keep;
    const float f = (float)(1.0 - (double)c * 0.01);
    const float omb = 1.0f - beta;
    state[2] = state[2] * beta;
    state[0] = state[0] * beta + (mn * f) * omb;
    state[1] = state[1] * beta + (mx * f) * omb;
    return keep;
}
