// api.cu -- error text, device queries and version of the nnmd_b200 C ABI.
#include <stdarg.h>

#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: ranges show up in nsys / ncu --nvtx when a tool is attached

#include "common.cuh"

namespace nnmd {

thread_local char g_err[512] = "";

int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
    return n;
}

std::atomic<long long> g_launches{0};

// Per-class event pairs recorded on the launching stream while profiling is enabled.
static bool g_prof_on = false;
static bool g_nvtx_on = false;
static const char *const g_class_names[PROF_NCLASS] = {"nnmd:K1 neighbour list", "nnmd:K2 radial", "nnmd:K3 angular", "nnmd:K5 normalise",
                                                       "nnmd:K4c contraction", "nnmd:K6 atomic network", "nnmd:K0 wrap", "nnmd:K4e edge gather"};
struct ProfRec { int cls; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t prof_event() {
    cudaEvent_t e;
    if (!g_pool.empty()) { e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int cls, cudaStream_t st) {
    if (g_nvtx_on) nvtxRangePushA(g_class_names[cls]);
    if (!g_prof_on) return;
    ProfRec r{cls, prof_event(), prof_event()};
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
void prof_end(int cls, cudaStream_t st) {
    if (g_nvtx_on) nvtxRangePop();
    if (!g_prof_on) return;
    for (size_t i = g_prof.size(); i-- > 0;)
        if (g_prof[i].cls == cls) { cudaEventRecord(g_prof[i].b, st); return; }
}

}  // namespace nnmd

extern "C" long long nnmd_launch_count(void) { return nnmd::g_launches.load(); }

extern "C" int nnmd_nvtx_enable(int on) {
    nnmd::g_nvtx_on = on != 0;
    return NNMD_OK;
}

extern "C" int nnmd_profile_enable(int on) {
    nnmd::g_prof_on = on != 0;
    return NNMD_OK;
}

// Synchronises the device, then adds up the recorded spans per class: ms[c], count[c] for c < NNMD_PROF_NCLASS. Resets.
extern "C" int nnmd_profile_read(double *ms, long long *count) {
    using namespace nnmd;
    if (cudaDeviceSynchronize() != cudaSuccess) return set_err(NNMD_ERR_CUDA, "profile_read: synchronize failed");
    for (int c = 0; c < PROF_NCLASS; ++c) { if (ms) ms[c] = 0; if (count) count[c] = 0; }
    for (ProfRec &r : g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            if (ms) ms[r.cls] += t;
            if (count) count[r.cls] += 1;
        }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_prof.clear();
    return NNMD_OK;
}

extern "C" int nnmd_abi_version(void) { return 1; }
extern "C" const char *nnmd_last_error(void) { return nnmd::g_err; }
extern "C" int nnmd_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
