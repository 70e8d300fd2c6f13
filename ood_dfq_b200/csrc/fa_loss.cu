// Feature-alignment loss of the QAT step, all residual units in ONE kernel each way.
//
// Reference: Trainer.loss_fa (trainer_direct.py:325-330) over the maps Trainer.channel_attention leaves
// (trainer_direct.py:382-383, hooks :432-440):
//
//     A = F.normalize(E)                    E[n,c] = mean_hw x[n,c,hw]^2 of a residual body output, [N, C_l]
//     fa = lam * sum_l mean_{n,c} (A_student_l - A_teacher_l)^2
//
// The energies E already come out of the fused residual tail / the channel-energy kernel.  What is left is tiny
// ([256, 64..512] per unit) but in eager PyTorch it is ~10 element-wise / reduction launches per unit and pass
// forward and as many again backward (norm, clamp, expand, div, sub, pow, mean, add and their tape): ~360 launches
// per iteration, 17 % of the device time of the 32x32 step and 4 % of the 224x224 one
// (profiles/r2_step_share_cifar.txt).  Here: one CTA per (unit, image) row normalises both rows and sums the squared
// difference; the last CTA folds the row sums in unit / row order (deterministic, no atomics).  Backward: the same
// grid writes dL/dE for student and teacher rows from the closed form of the normalisation's derivative.
//
// Arithmetic: sums over C and over rows are accumulated in fp64 and rounded once to fp32 (ATen: fp32 trees); the
// quotient E / max(||E||, 1e-12) is an fp32 IEEE division as in F.normalize.  Agreement with the ATen chain: ~1e-7
// relative (tests/test_gpu_fa_loss.py holds 1e-5).  Launch-latency-bound, not a roofline kernel.
#include "bn_geom.cuh"

namespace oodfq {

constexpr int kFaMaxLayers = 32;
constexpr int kFaThreads = 128;
constexpr float kFaEps = 1e-12f;          // F.normalize's default eps

struct FaTable {
    const float* es[kFaMaxLayers];
    const float* et[kFaMaxLayers];
    float* ges[kFaMaxLayers];
    float* get[kFaMaxLayers];
    int C[kFaMaxLayers];
    int L, N;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                        // red may still be read by the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kFaThreads / 32; ++w) t += red[w];
    return t;
}

// row norms as F.normalize forms them: ||E||_2 rounded to fp32, clamped from below
__device__ __forceinline__ void row_norms(const float* es, const float* et, int C, double* red, float& ms, float& mt) {
    double s = 0.0, t = 0.0;
    for (int c = threadIdx.x; c < C; c += kFaThreads) {
        const double a = (double)__ldg(es + c), b = (double)__ldg(et + c);
        s += a * a;
        t += b * b;
    }
    s = block_sum(s, red);
    t = block_sum(t, red);
    ms = fmaxf((float)sqrt(s), kFaEps);
    mt = fmaxf((float)sqrt(t), kFaEps);
}

__global__ void __launch_bounds__(kFaThreads)
fa_loss_fwd_kernel(const FaTable T, float lam, double* __restrict__ rowsum, float* __restrict__ loss, int* ticket) {
    __shared__ double red[kFaThreads / 32];
    __shared__ int s_last;
    const int l = blockIdx.x / T.N, n = blockIdx.x % T.N, C = T.C[l];
    const float* es = T.es[l] + (long long)n * C;
    const float* et = T.et[l] + (long long)n * C;
    float ms, mt;
    row_norms(es, et, C, red, ms, mt);
    double acc = 0.0;
    for (int c = threadIdx.x; c < C; c += kFaThreads) {
        const float d = __fsub_rn(__fdiv_rn(__ldg(es + c), ms), __fdiv_rn(__ldg(et + c), mt));
        acc += (double)d * (double)d;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        rowsum[blockIdx.x] = acc;
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: per-unit means in row order, then the reference's running sum over units and the factor lam
    float fa = 0.0f;
    for (int u = 0; u < T.L; ++u) {
        double t = 0.0;
        for (int r = threadIdx.x; r < T.N; r += kFaThreads) t += __ldcg(rowsum + (long long)u * T.N + r);
        t = block_sum(t, red);
        fa = __fadd_rn(fa, (float)(t / ((double)T.N * (double)T.C[u])));
    }
    if (threadIdx.x == 0) {
        loss[0] = __fmul_rn(lam, fa);
        *ticket = 0;
    }
}

// dL/dA_s = g * lam * 2 (A_s - A_t) / (N C), dL/dA_t = -dL/dA_s;  A = E / m with m = max(||E||, eps):
//   ||E|| > eps :  dE = (dA - A * <A, dA>) / m          ||E|| <= eps :  dE = dA / m   (the clamp passes no gradient)
__global__ void __launch_bounds__(kFaThreads)
fa_loss_bwd_kernel(const FaTable T, float lam, const float* __restrict__ gscale) {
    __shared__ double red[kFaThreads / 32];
    const int l = blockIdx.x / T.N, n = blockIdx.x % T.N, C = T.C[l];
    const long long off = (long long)n * C;
    const float* es = T.es[l] + off;
    const float* et = T.et[l] + off;
    float ms, mt;
    row_norms(es, et, C, red, ms, mt);
    const double k = (double)(gscale ? __ldg(gscale) : 1.0f) * (double)lam * 2.0 / ((double)T.N * (double)C);
    double ds = 0.0, dt = 0.0;                  // <A_s, dA_s>, <A_t, dA_t>
    for (int c = threadIdx.x; c < C; c += kFaThreads) {
        const float as = __fdiv_rn(__ldg(es + c), ms), at = __fdiv_rn(__ldg(et + c), mt);
        const double da = k * (double)__fsub_rn(as, at);
        ds += (double)as * da;
        dt -= (double)at * da;
    }
    ds = block_sum(ds, red);
    dt = block_sum(dt, red);
    const bool cs = !(ms > kFaEps), ct = !(mt > kFaEps);       // clamped rows
    for (int c = threadIdx.x; c < C; c += kFaThreads) {
        const float as = __fdiv_rn(__ldg(es + c), ms), at = __fdiv_rn(__ldg(et + c), mt);
        const double da = k * (double)__fsub_rn(as, at);
        if (T.ges[l]) T.ges[l][off + c] = (float)((da - (cs ? 0.0 : (double)as * ds)) / (double)ms);
        if (T.get[l]) T.get[l][off + c] = (float)((-da - (ct ? 0.0 : (double)at * dt)) / (double)mt);
    }
}

static int fill_table(FaTable& T, const float* const* es, const float* const* et, float* const* ges, float* const* get,
                      const int* channels, int L, int N) {
    if (!es || !et || !channels || L < 1 || L > kFaMaxLayers || N < 1) return OODFQ_EINVAL;
    T.L = L;
    T.N = N;
    for (int l = 0; l < L; ++l) {
        if (!es[l] || !et[l] || channels[l] < 1) return OODFQ_EINVAL;
        T.es[l] = es[l];
        T.et[l] = et[l];
        T.ges[l] = ges ? ges[l] : nullptr;
        T.get[l] = get ? get[l] : nullptr;
        T.C[l] = channels[l];
    }
    return OODFQ_OK;
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_fa_loss_max_layers(void) { return kFaMaxLayers; }

extern "C" int oodfq_fa_loss_forward(const float* const* e_student, const float* const* e_teacher, const int* channels,
                                     int L, int N, float lam, float* loss, double* row_scratch, void* workspace,
                                     oodfq_stream_t stream) {
    FaTable T;
    if (!loss || !row_scratch || !workspace || fill_table(T, e_student, e_teacher, nullptr, nullptr, channels, L, N) != OODFQ_OK)
        return fail(OODFQ_EINVAL, "fa_loss_forward: needs 1..%d units of [N, C] energies, a loss scalar, L*N doubles of scratch", kFaMaxLayers);
    Workspace* ws = reinterpret_cast<Workspace*>(workspace);
    fa_loss_fwd_kernel<<<(unsigned)(L * N), kFaThreads, 0, (cudaStream_t)stream>>>(T, lam, row_scratch, loss, &ws->ticket[1]);
    count_launch();
    return check_launch("fa_loss_forward");
}

extern "C" int oodfq_fa_loss_backward(const float* const* e_student, const float* const* e_teacher, const int* channels,
                                      int L, int N, float lam, const float* grad_loss, float* const* grad_student,
                                      float* const* grad_teacher, oodfq_stream_t stream) {
    FaTable T;
    if ((!grad_student && !grad_teacher) ||
        fill_table(T, e_student, e_teacher, grad_student, grad_teacher, channels, L, N) != OODFQ_OK)
        return fail(OODFQ_EINVAL, "fa_loss_backward: needs 1..%d units of [N, C] energies and at least one gradient table", kFaMaxLayers);
    fa_loss_bwd_kernel<<<(unsigned)(L * N), kFaThreads, 0, (cudaStream_t)stream>>>(T, lam, grad_loss);
    count_launch();
    return check_launch("fa_loss_backward");
}
