// Library bookkeeping: ABI version, last-error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace oodfq {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OODFQ_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return OODFQ_OK;
}

}  // namespace oodfq

extern "C" {

int oodfq_abi_version(void) { return OODFQ_ABI_VERSION; }
const char* oodfq_last_error(void) { return oodfq::g_err; }
unsigned long long oodfq_launch_count(void) { return oodfq::g_launches.load(); }
void oodfq_reset_launch_count(void) { oodfq::g_launches.store(0); }
size_t oodfq_workspace_bytes(void) { return oodfq::kWorkspaceBytes; }

}  // extern "C"
