// Grid-level fold of the per-CTA (dW, dB) partials of the channels_last reducing kernels -- immediately (one small
// launch behind the producing kernel) or DEFERRED: between oodfq_defer_folds_begin and _end every producing kernel
// writes its partials into its own region of a caller-provided arena and only leaves a note; _flush / _end then fold
// everything noted with ONE launch.  A backward sweep of the QAT step has 17-19 such reductions (every BatchNorm's
// weight / bias gradient), each followed by a ~4 us fold launch that nothing reads before the optimiser does: 34-38
// launches per iteration, 2.4 % of the 32x32 step (profiles/r2_step_share_cifar.txt).  The fold arithmetic is the same
// function either way (fold_partials: one warp per channel, fixed order), so the results are bit-identical.
//
// The caller's side of the bargain: an output of a deferred reduction holds garbage until the flush, so nothing may
// read it in between (step.QATStep flushes at the end of each autograd sweep, before the gradients are moved), and
// the output buffers must stay allocated until then (ops.py keeps them).
#include <mutex>
#include <vector>

#include "bn_geom.cuh"

namespace oodfq {

struct FoldDesc {
    const double* partial;   // [nparts][C][2]
    void* out;               // [2C] floats (parameter gradients) or doubles (BN-input statistics)
    int C, nparts, out_is_double;
};
constexpr int kFoldBatchMax = 96;
struct FoldBatch {
    FoldDesc d[kFoldBatchMax];
    int first_cta[kFoldBatchMax + 1];
    int n;
};

__global__ void __launch_bounds__(kBThreads) bn_fold_multi_kernel(const __grid_constant__ FoldBatch B) {
    int i = 0;
    while (i + 1 < B.n && (int)blockIdx.x >= B.first_cta[i + 1]) ++i;
    const int c = ((int)blockIdx.x - B.first_cta[i]) * (kBThreads / 32) + (threadIdx.x >> 5);
    if (c >= B.d[i].C) return;
    if (B.d[i].out_is_double) fold_partials(B.d[i].partial, B.d[i].C, c, B.d[i].nparts, threadIdx.x & 31, static_cast<double*>(B.d[i].out));
    else fold_partials(B.d[i].partial, B.d[i].C, c, B.d[i].nparts, threadIdx.x & 31, static_cast<float*>(B.d[i].out));
}

namespace {
std::mutex g_mu;
bool g_on = false;
char* g_arena = nullptr;
size_t g_bytes = 0, g_used = 0;
std::vector<FoldDesc> g_pending;

int flush_locked(cudaStream_t st) {
    size_t at = 0;
    while (at < g_pending.size()) {
        FoldBatch B;
        B.n = 0;
        int ctas = 0;
        while (at < g_pending.size() && B.n < kFoldBatchMax) {
            B.d[B.n] = g_pending[at++];
            B.first_cta[B.n] = ctas;
            ctas += (B.d[B.n].C + kBThreads / 32 - 1) / (kBThreads / 32);
            ++B.n;
        }
        B.first_cta[B.n] = ctas;
        bn_fold_multi_kernel<<<(unsigned)ctas, kBThreads, 0, st>>>(B);
        count_launch();
    }
    g_pending.clear();
    g_used = 0;
    return check_launch("fold (deferred)");
}
}  // namespace

double* fold_target(double* ws_partial, int C, int nparts) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_on) return ws_partial;
    const size_t need = (((size_t)nparts * C * 2 * sizeof(double)) + 255) & ~(size_t)255;
    if (g_used + need > g_bytes) return ws_partial;          // arena full: this one folds immediately
    double* p = reinterpret_cast<double*>(g_arena + g_used);
    g_used += need;
    return p;
}

template <typename OutT>
static int fold_finish_t(double* target, double* ws_partial, int C, int nparts, OutT* out, cudaStream_t st) {
    if (target == ws_partial) {
        bn_nhwc_fold_kernel<OutT><<<(C + kBThreads / 32 - 1) / (kBThreads / 32), kBThreads, 0, st>>>(ws_partial, C, nparts, out);
        count_launch();
        return check_launch("fold");
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_pending.push_back(FoldDesc{target, out, C, nparts, sizeof(OutT) == sizeof(double) ? 1 : 0});
    return OODFQ_OK;
}

int fold_finish(double* target, double* ws_partial, int C, int nparts, float* out, cudaStream_t st) {
    return fold_finish_t(target, ws_partial, C, nparts, out, st);
}
int fold_finish(double* target, double* ws_partial, int C, int nparts, double* out, cudaStream_t st) {
    return fold_finish_t(target, ws_partial, C, nparts, out, st);
}

}  // namespace oodfq

using namespace oodfq;

extern "C" int oodfq_defer_folds_begin(void* arena, size_t bytes) {
    if (!arena || bytes < 4096 || (reinterpret_cast<uintptr_t>(arena) & 255u))
        return fail(OODFQ_EINVAL, "defer_folds_begin: needs a 256-byte aligned device arena of at least 4 KB");
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_on) return fail(OODFQ_EINVAL, "defer_folds_begin: already deferring (flush / end first)");
    g_on = true;
    g_arena = static_cast<char*>(arena);
    g_bytes = bytes;
    g_used = 0;
    g_pending.clear();
    return OODFQ_OK;
}

extern "C" int oodfq_defer_folds_flush(oodfq_stream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_on) return OODFQ_OK;
    return flush_locked((cudaStream_t)stream);
}

extern "C" int oodfq_defer_folds_end(oodfq_stream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_on) return OODFQ_OK;
    const int rc = flush_locked((cudaStream_t)stream);
    g_on = false;
    g_arena = nullptr;
    g_bytes = 0;
    return rc;
}

extern "C" int oodfq_defer_folds_pending(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    return (int)g_pending.size();
}
