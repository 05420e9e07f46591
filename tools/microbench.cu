// microbench.cu -- measured FP32 / packed-FP32 / SFU peaks of the GPU (denominators for the compute-bound kernels).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu && ./tools/microbench
// Prints one JSON object: warp-instructions per clock per SM and lane-ops per second for FFMA, FFMA2 (fma.rn.f32x2),
// MUFU.EX2, and two mixes shaped like the inner loop of the angular kernel.
#include <cuda_runtime.h>
#include <stdio.h>

#define ITERS 4096

__global__ void k_ffma(float *out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b) {
    float2 x[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with three VARYING register-pair operands (the form the angular inner loop uses for q*t + acc)
__global__ void k_ffma2_3op(float *out, float a, float b) {
    float2 x[6], y[6], z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        y[i] = make_float2(a + i * 1e-4f, a - i * 1e-4f);
        z[i] = make_float2(b + i * 1e-5f, b - i * 1e-5f);
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = __ffma2_rn(y[i], z[(i + 1) % 6], x[i]);
#pragma unroll
        for (int i = 0; i < 6; ++i) y[i] = __ffma2_rn(z[i], x[(i + 2) % 6], y[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) s += x[i].x + x[i].y + y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scalar FFMA with three varying operands, same dataflow
__global__ void k_ffma_3op(float *out, float a, float b) {
    float x[12], y[12], z[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = a + i * 1e-4f; z[i] = b + i * 1e-5f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = fmaf(y[i], z[(i + 1) % 12], x[i]);
#pragma unroll
        for (int i = 0; i < 12; ++i) y[i] = fmaf(z[i], x[(i + 2) % 12], y[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__global__ void k_mufu(float *out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i * a;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = ex2(x[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix A: 7 scalar FFMA-pipe ops per MUFU (the scalar inner loop: 44 FP32 + 6 MUFU)
__global__ void k_mix_scalar(float *out, float a, float b) {
    float x[14], m[2];
#pragma unroll
    for (int i = 0; i < 14; ++i) x[i] = threadIdx.x * 1e-3f + i;
    m[0] = a; m[1] = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 14; ++i) x[i] = fmaf(x[i], a, b);
        m[0] = ex2(m[0]); m[1] = ex2(m[1]);
    }
    float s = m[0] + m[1];
#pragma unroll
    for (int i = 0; i < 14; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix B: the same arithmetic with packed FMAs: 7 FFMA2 per 2 MUFU
__global__ void k_mix_packed(float *out, float a, float b) {
    float2 x[7];
    float m[2];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    m[0] = a; m[1] = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
        m[0] = ex2(m[0]); m[1] = ex2(m[1]);
    }
    float s = m[0] + m[1];
#pragma unroll
    for (int i = 0; i < 7; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float t; cudaEventElapsedTime(&t, a, b);
        if (t < best) best = t;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double warps = double(blocks) * threads / 32.0;
    const double t_ffma = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 1e-4f); });
    const double t_ffma2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 1e-4f); });
    const double t_mufu = time_ms([&] { k_mufu<<<blocks, threads>>>(out, -0.5f); });
    const double t_mixs = time_ms([&] { k_mix_scalar<<<blocks, threads>>>(out, 0.9999f, -0.5f); });
    const double t_mixp = time_ms([&] { k_mix_packed<<<blocks, threads>>>(out, 0.9999f, -0.5f); });
    const double t_f2_3 = time_ms([&] { k_ffma2_3op<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
    const double t_f1_3 = time_ms([&] { k_ffma_3op<<<blocks, threads>>>(out, 0.999f, 1e-3f); });
    auto lane_ops = [&](double per_iter, double ms) { return warps * 32.0 * ITERS * per_iter / (ms * 1e-3); };
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d,\n", p.name, sms, clk_khz);
    printf(" \"ffma_lane_ops_per_s\": %.4e, \"ffma_ms\": %.4f,\n", lane_ops(16, t_ffma), t_ffma);
    printf(" \"ffma2_fma_lane_ops_per_s\": %.4e, \"ffma2_ms\": %.4f,\n", lane_ops(16, t_ffma2), t_ffma2);
    printf(" \"ffma2_3op_fma_lane_ops_per_s\": %.4e, \"ffma_3op_lane_ops_per_s\": %.4e,\n", lane_ops(24, t_f2_3), lane_ops(24, t_f1_3));
    printf(" \"mufu_ex2_lane_ops_per_s\": %.4e, \"mufu_ms\": %.4f,\n", lane_ops(8, t_mufu), t_mufu);
    printf(" \"mix_scalar_14ffma_2mufu_ms\": %.4f, \"mix_packed_7ffma2_2mufu_ms\": %.4f,\n", t_mixs, t_mixp);
    printf(" \"fp32_tflops_ffma\": %.2f, \"fp32_tflops_ffma2\": %.2f}\n", 2e-12 * lane_ops(16, t_ffma), 2e-12 * lane_ops(16, t_ffma2));
    return 0;
}
