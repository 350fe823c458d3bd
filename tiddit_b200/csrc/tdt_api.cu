// tdt_api.cu -- error reporting, launch counter, version (libtdt_b200.so, sm_100a).
#include <stdarg.h>
#include <stdlib.h>

#include "tdt_common.cuh"

namespace tdt {

thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

struct ProfStage {
    const char *name;
    cudaEvent_t a, b;
};
static bool g_prof_on = false;
static bool g_prof_detail = false;   // TDT_PROF_DETAIL=1: one entry per kernel launch instead of per stage
static int g_prof_n = 0;
constexpr int PROF_MAX = 2048;
static ProfStage g_prof[PROF_MAX];

void prof_kernel_begin(const char *name, cudaStream_t st) {
    if (!g_prof_on || !g_prof_detail || g_prof_n >= PROF_MAX) return;
    ProfStage &s = g_prof[g_prof_n];
    s.name = name;
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    cudaEventRecord(s.a, st);
}

void prof_kernel_end(cudaStream_t st) {
    if (!g_prof_on || !g_prof_detail || g_prof_n >= PROF_MAX) return;
    cudaEventRecord(g_prof[g_prof_n].b, st);
    g_prof_n++;
}

void prof_stage_begin(const char *name, cudaStream_t st) {
    if (!g_prof_on || g_prof_detail || g_prof_n >= PROF_MAX) return;
    ProfStage &s = g_prof[g_prof_n];
    s.name = name;
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    cudaEventRecord(s.a, st);
}

void prof_stage_end(cudaStream_t st) {
    if (!g_prof_on || g_prof_detail || g_prof_n >= PROF_MAX) return;
    cudaEventRecord(g_prof[g_prof_n].b, st);
    g_prof_n++;
}

}  // namespace tdt

extern "C" {

void tdt_profile_begin(void) {
    const char *d = getenv("TDT_PROF_DETAIL");
    tdt::g_prof_detail = d && d[0] == '1';
    tdt::g_prof_on = true;
    tdt::g_prof_n = 0;
}

// writes "stage=milliseconds\n" lines (stages in call order) and switches profiling off; returns the
// number of bytes written (without the terminating 0) or < 0 on error
int tdt_profile_end(char *out, size_t cap) {
    using namespace tdt;
    g_prof_on = false;
    size_t off = 0;
    for (int i = 0; i < g_prof_n; i++) {
        float ms = 0.f;
        TDT_CUDA(cudaEventSynchronize(g_prof[i].b));
        TDT_CUDA(cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b));
        cudaEventDestroy(g_prof[i].a);
        cudaEventDestroy(g_prof[i].b);
        if (out && off < cap) {
            int w = snprintf(out + off, cap - off, "%s=%.6f\n", g_prof[i].name, ms);
            if (w > 0) off += (size_t)w < cap - off ? (size_t)w : cap - off - 1;
        }
    }
    g_prof_n = 0;
    return (int)off;
}

int tdt_version(void) { return 100; }  // 0.1.0

const char *tdt_last_error(void) { return tdt::g_err; }

int64_t tdt_launch_count(void) { return tdt::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
