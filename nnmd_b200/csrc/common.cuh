// common.cuh -- shared helpers of the nnmd_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nnmd_b200.h"

namespace nnmd {

// ---- error text -------------------------------------------------------------------------
extern thread_local char g_err[512];
int set_err(int code, const char *fmt, ...);

#define NNMD_CUDA_CHECK(expr)                                                                     \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ::nnmd::set_err(NNMD_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

#define NNMD_LAUNCH_CHECK(name)                                                                   \
    do {                                                                                          \
        ::nnmd::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess)                                                                   \
            return ::nnmd::set_err(NNMD_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

int sm_count();

// Tuning / profiling knobs are read from the environment only by libraries built with -DNNMD_TUNING
// (tools/build_variant.sh); the product library ignores them, so a stray exported variable cannot change results.
#ifdef NNMD_TUNING
inline const char *tuning_env(const char *name) { return getenv(name); }
#else
inline const char *tuning_env(const char *) { return nullptr; }
#endif

// ---- launch counter and optional per-kernel-class CUDA-event timing (bench.py's roofline leg) ----------
extern std::atomic<long long> g_launches;
enum ProfClass { PROF_NLIST = 0, PROF_RADIAL = 1, PROF_ANGULAR = 2, PROF_NORMALIZE = 3, PROF_CONTRACT = 4, PROF_MLP = 5,
                 PROF_WRAP = 6, PROF_GATHER = 7, PROF_NCLASS = 8 };
void prof_begin(int cls, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
struct ProfScope {
    int cls; cudaStream_t st;
    ProfScope(int c, cudaStream_t s) : cls(c), st(s) { prof_begin(cls, st); }
    ~ProfScope() { prof_end(cls, st); }
};

// ---- 4-wide position type: fp32 -> float4 (one 16 B load), fp64 -> double4 (two 16 B loads)
template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename T> __device__ __forceinline__ void load_pos(const T *posw, int64_t i, T &x, T &y, T &z);
template <> __device__ __forceinline__ void load_pos<float>(const float *posw, int64_t i, float &x, float &y, float &z) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(posw) + i);
    x = v.x; y = v.y; z = v.z;
}
template <> __device__ __forceinline__ void load_pos<double>(const double *posw, int64_t i, double &x, double &y, double &z) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(posw) + 2 * i);
    const double2 b = __ldg(reinterpret_cast<const double2 *>(posw) + 2 * i + 1);
    x = a.x; y = a.y; z = b.x;
}

// position + the 4th component (species index of the atom as a real number, 0 when species are not used)
template <typename T> __device__ __forceinline__ void load_pos4(const T *posw, int64_t i, T &x, T &y, T &z, T &w);
template <> __device__ __forceinline__ void load_pos4<float>(const float *posw, int64_t i, float &x, float &y, float &z, float &w) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(posw) + i);
    x = v.x; y = v.y; z = v.z; w = v.w;
}
template <> __device__ __forceinline__ void load_pos4<double>(const double *posw, int64_t i, double &x, double &y, double &z, double &w) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(posw) + 2 * i);
    const double2 b = __ldg(reinterpret_cast<const double2 *>(posw) + 2 * i + 1);
    x = a.x; y = a.y; z = b.x; w = b.y;
}

// ---- math in the working precision ----------------------------------------------------------
// "acc" = accurate library routine (used once per neighbour / per triangle),
// "fast" = the SFU path used only inside the per-function inner loop of the fp32 kernels.
template <typename T> struct M;
template <> struct M<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ void sincospi_(float x, float *s, float *c) { sincospif(x, s, c); }
    static __device__ __forceinline__ float cospi_(float x) { return cospif(x); }
    static __device__ __forceinline__ float log2_(float x) { return log2f(x); }
    static __device__ __forceinline__ float exp2_fast(float x) {
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float log2_fast(float x) {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    // SFU forms used once per triangle in the fp32 kernels (each within a few ulp; error budget in DESIGN.md)
    static __device__ __forceinline__ float rsqrt_fast(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ float cospi_fast(float x) { return __cosf(x * 3.14159265358979323846f); }
    static __device__ __forceinline__ float exp_fast(float x) { return exp2_fast(x * 1.4426950408889634f); }
    static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
    static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float floor_(float x) { return floorf(x); }
    static __device__ __forceinline__ float big() { return 1e30f; }
    static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
};
template <> struct M<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ void sincospi_(double x, double *s, double *c) { sincospi(x, s, c); }
    static __device__ __forceinline__ double cospi_(double x) { return cospi(x); }
    static __device__ __forceinline__ double log2_(double x) { return log2(x); }
    static __device__ __forceinline__ double exp2_fast(double x) { return exp2(x); }
    static __device__ __forceinline__ double log2_fast(double x) { return log2(x); }
    static __device__ __forceinline__ double rsqrt_fast(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double cospi_fast(double x) { return cospi(x); }
    static __device__ __forceinline__ double exp_fast(double x) { return exp(x); }
    static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
    static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double floor_(double x) { return floor(x); }
    static __device__ __forceinline__ double big() { return 1e300; }
    static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
};

// |u| with a fixed operation order and no FMA contraction: every kernel that decides "is this neighbour inside the
// cutoff" gets bit-identical answers (the edge-centric kernels rely on it).
template <typename T> __device__ __forceinline__ T norm3(T x, T y, T z);
template <> __device__ __forceinline__ float norm3<float>(float x, float y, float z) {
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
template <> __device__ __forceinline__ double norm3<double>(double x, double y, double z) {
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace nnmd
